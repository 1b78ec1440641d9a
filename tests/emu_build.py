"""Builds ONE shared library with everything the emulation tests need (tests/emu/): the library's own sources compiled
with g++ -DDKT_EMU (translation units in parallel) plus the test harnesses.  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "dendro-kt_b200", "csrc")
# DKT_EMU_CFLAGS="-DDKT_FAM_ALIAS=1 ..." emulates a build variant (its objects live in a directory of their own)
EXTRA = os.environ.get("DKT_EMU_CFLAGS", "").split()
BUILD = os.path.join(EMU, "_build" + ("_" + "".join(ch if ch.isalnum() else "_" for ch in "".join(EXTRA)) if EXTRA else ""))
LIB = os.path.join(BUILD, "libdkt_emu_all.so")
SOURCES = [os.path.join(CSRC, f) for f in ("dkt_build.cu", "dkt_chunks.cu", "dkt_family.cu", "dkt_matvec.cu", "dkt_dist.cu", "dkt_solve.cu", "dkt_tree.cu", "dkt_sfc.cpp")] + \
          [os.path.join(EMU, f) for f in ("cuda_emu.cpp", "emu_common.cpp", "emu_harness.cpp", "emu_full.cpp", "emu_dist.cpp", "emu_p2p.cpp")]
HEADERS = [os.path.join(EMU, "cuda_emu.h"), os.path.join(CSRC, "dkt_internal.h"), os.path.join(CSRC, "dkt_chunks.h"), os.path.join(CSRC, "dkt_p2p.cuh"),
           os.path.join(ROOT, "include", "dkt.h")]


def build():
    os.makedirs(BUILD, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in HEADERS)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(BUILD, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            cmd = ["g++", "-std=c++17", "-O1", "-g", "-DDKT_EMU", *EXTRA, "-Wno-unknown-pragmas", "-I" + EMU, "-I" + CSRC, "-fPIC", "-x", "c++", "-c",
                   src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append("%s\n%s" % (src, out))
    if failed:
        raise RuntimeError("emulation build failed:\n" + "\n".join(failed))
    if procs or not os.path.exists(LIB):
        subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB
