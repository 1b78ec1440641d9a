#include "mathUtils.h"
