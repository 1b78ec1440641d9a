/* dkt.h - C ABI of libdkt.so: the B200-native matrix-free FE matvec on adaptive k-D SFC trees.
 *
 * This is the drop-in boundary for ONE path of paralab/Dendro-KT: everything between
 * `ot::DA<dim>` construction from a balanced linear tree and `feMatrix<LeafT,dim>::matVec`.
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository).  Plain pointers and sizes only; no exceptions cross this boundary;
 * every function returns an int status (DKT_OK == 0) and dkt_last_error() describes the
 * most recent failure on the calling thread.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a).
 */
#ifndef DKT_H
#define DKT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DKT_OK 0
#define DKT_ERR_INVALID 1   /* bad argument                                              */
#define DKT_ERR_CUDA 2      /* CUDA runtime / driver failure (no device, OOM, ...)       */
#define DKT_ERR_UNSUPPORTED 3
#define DKT_ERR_UNDEFINED_TREE 4 /* class-U tree: the reference reads undefined values   */
#define DKT_ERR_NCCL 5

/* SFC used for the element order and inside the node order (CMake option HILBERT_ORDERING,
 * src/KDhcurvedata.cpp:24-59).  The reference's default build is Morton. */
#define DKT_SFC_MORTON 0
#define DKT_SFC_HILBERT 1

/* dkt_da_create flags */
#define DKT_ELEMS_ON_DEVICE 1u  /* elem_xyz / elem_lev are device pointers                  */
#define DKT_ELEMS_PRESORTED 2u  /* elements already in tree order (skip the SFC sort)        */
#define DKT_ALLOW_UNDEFINED 4u  /* build class-U trees anyway (intended semantics, unpinned) */
#define DKT_DIST_DRYRUN 8u      /* dkt_da_create_dist: build rank's tables but no communicator (nccl_id may be
                                   NULL); for inspecting the partition on one GPU - dkt_matvec then fails */

/* dkt_matvec flags */
#define DKT_VEC_HOST 0u    /* in/out are host pointers: H2D, matvec, D2H                    */
#define DKT_VEC_DEVICE 1u  /* in/out are device pointers on the DA's device                */
#define DKT_NO_Q1_MASK 2u  /* mathematically consistent transpose (NOT the reference's)    */
#define DKT_MV_FLAT 4u     /* use the flat gather/atomic kernels instead of the chunked ones */
#define DKT_VEC_GHOSTED 16u /* partitioned DA, device vectors: in/out hold n_nodes + n_ghost_nodes entries
                              ([owned | ghosts], like the reference's ghosted vectors) and are used in place */
#define DKT_MV_NO_FASTPATH 8u /* apply kref as a dense matrix even if it has Walsh-Hadamard diagonal form */

/* tree classes */
#define DKT_CLASS_A 0 /* no hanging nodes                                                  */
#define DKT_CLASS_B 1 /* hanging nodes, none on the domain boundary                        */
#define DKT_CLASS_P 2 /* boundary hanging nodes kept as DOFs -> phantom child elements      */
#define DKT_CLASS_U 3 /* reference reads undefined parent values (FEM/include/matvec.h:439) */

/* elemental operator kinds */
#define DKT_OP_IDENTITY 0 /* out = in  (test/testMatvec.cpp:177-198)                        */
#define DKT_OP_DENSE 1    /* out = scale * 2^(-alpha*level) * kref * in                     */
#define DKT_OP_KRON 2     /* sum-factorised: out = scale * 2^(-alpha*level) * sum_t (A[t][dim-1] x .. x A[t][0]) in,
                             the form of HeatMat / HeatVec (FEM/examples/src/heatMat.cpp:46-117, FEM/src/tensor.cpp:19-107) */
#define DKT_KRON_MAX_TERMS 5

typedef struct dkt_da dkt_da;

/* Device form of the user's `elementalMatVec(const VECType*in, VECType*out, double*coords,
 * double scale)` (FEM/include/feMatrix.h:56).  The host callback cannot run per element on the
 * GPU; a leaf class maps to one of these (and the C++ layer verifies the mapping against the
 * callback on sample elements).  For axis-aligned cells K_e depends only on the level. */
typedef struct dkt_op
{
  int kind;            /* DKT_OP_*                                                           */
  const double *kref;  /* host pointer, N*N row-major: out[i] = sum_j kref[i*N+j]*in[j]      */
  double alpha;        /* level exponent: K_e = scale * 2^(-alpha*L) * kref                 */
  int dirichlet;       /* 1: zero domain-boundary entries of input and output, like
                          HeatMat::preMatVec/postMatVec (FEM/examples/src/heatMat.cpp:120-139) */
  int terms;           /* DKT_OP_KRON: number of Kronecker terms (<= DKT_KRON_MAX_TERMS); kref then holds terms * dim matrices of
                          M x M doubles (M = order + 1), term-major, axis 0 first: A[t][d][k*M + j] takes input index k to output
                          index j along axis d - the layout of the reference's 1-D operators (FEM/include/refel.h).  Order 2
                          runs them as axis passes in registers; order 1 and the flat kernels expand them to kref.           */
} dkt_op;

typedef struct dkt_sizes
{
  uint64_t n_elem;       /* tree elements (leaves)                                           */
  uint64_t n_mv_elem;    /* elements the matvec visits (phantom children included)           */
  uint64_t n_nodes;      /* CG nodes == DA::getTotalNodalSz() on one rank                   */
  uint64_t n_boundary;   /* DA::getBoundaryNodeIndices().size()                              */
  uint64_t n_hanging;    /* visited elements with at least one hanging lattice node          */
  uint64_t n_split;      /* tree elements replaced by phantom children (class P)             */
  uint64_t alg_bytes;    /* algorithmic bytes of one matvec (SURVEY.md §8d formula)          */
  int nodes_per_elem;    /* (order+1)^dim                                                    */
  int tree_class;        /* DKT_CLASS_*                                                      */
  int finest_level;
  int n_ranks;           /* 1 for a single-rank DA                                           */
  uint64_t n_global_nodes; /* partitioned DA: CG nodes of the whole tree (n_nodes = owned ones) */
  uint64_t n_ghost_nodes;  /* partitioned DA: nodes this rank touches but does not own          */
  uint64_t n_global_elem;
} dkt_sizes;

const char *dkt_last_error(void);
const char *dkt_version(void);

/* SFC tables in the reference's layout (include/hcurvedata.h:49-59): rotations[r*2*2^dim + i]
 * = rot_perm (i < 2^dim) | rot_inv, hilbert_table[r*2^dim + child_m] = child rotation.
 * Generated from first principles (no table data is taken from the reference); the state
 * numbering differs from KDhcurvedata_DATA.cpp but the curve is the same.  Returns the number
 * of rotations; pass NULL outputs to query it.  Replaces _InitializeHcurve(dim). */
int dkt_sfc_tables(int dim, int sfc_mode, char *rotations, int *hilbert_table);

/* Replaces ot::DA<dim>::DA(const TreeNode*inTree, nEle, comm, order, ...) (include/oda.h:150,
 * src/oda.cpp:46-151) for a single rank.  inTree is passed as anchors elem_xyz[n_elem*dim]
 * (units of 2^-max_depth, i.e. TreeNode::getX(d)) and levels elem_lev[n_elem]; it must be a
 * 2:1-balanced complete linear tree.  ip0/ip1: RefElement::getIMChild0()/getIMChild1()
 * ((order+1)^2 doubles each, FEM/include/refel.h:187-188) or NULL for exact interpolation.
 * Builds on the current CUDA device: tree order (SFC_Tree::locTreeSort), the CG node set and
 * its order (SFC_NodeSort::dist_countCGNodes), element->node and hanging-node tables. */
int dkt_da_create(int dim, int order, int max_depth, int sfc_mode, const uint32_t *elem_xyz, const uint8_t *elem_lev,
                  uint64_t n_elem, const double *ip0, const double *ip1, unsigned flags, dkt_da **out);
int dkt_da_destroy(dkt_da *da);

/* Trees from points on the GPU.  Replaces SFC_Tree<T,dim>::distTreeConstruction (balance == 0) and distTreeBalancing
 * (balance != 0) of the reference on one rank (src/tsort.cpp:566-716, 775-877; include/tsort.h:228-304): pts_xyz holds
 * n_pts points, dim integer coordinates each in [0, 2^max_depth) (what TreeNode(coords, m_uiMaxDepth) holds); a region
 * is split while it contains more than max_pts_per_region points and is coarser than max_depth; balancing adds what
 * propagateNeighbours + the second construction add.  The result is the reference's leaf set in the reference's tree
 * order (sfc_mode), bit-exact.  flags: DKT_ELEMS_ON_DEVICE = pts_xyz (and the export targets) are device pointers.
 * dkt_tree_device_ptrs + dkt_da_create(..., DKT_ELEMS_ON_DEVICE | DKT_ELEMS_PRESORTED) builds a DA without a host
 * round trip.  Limits: dim * max_depth <= 56, fewer than 2^32 points and leaves. */
typedef struct dkt_tree dkt_tree;
int dkt_tree_from_points(int dim, int max_depth, int sfc_mode, const uint32_t *pts_xyz, uint64_t n_pts, uint64_t max_pts_per_region,
                         int balance, unsigned flags, dkt_tree **out);
int dkt_tree_size(const dkt_tree *t, uint64_t *n_elem, int *finest_level /* may be NULL */);
int dkt_tree_export(const dkt_tree *t, uint32_t *elem_xyz, uint8_t *elem_lev, unsigned flags);
int dkt_tree_device_ptrs(const dkt_tree *t, const uint32_t **elem_xyz, const uint8_t **elem_lev); /* owned by the tree */
int dkt_tree_destroy(dkt_tree *t);

/* Multi-GPU, one process per GPU.  Replaces the distributed DA of the reference for the matvec
 * path: SFC-contiguous element ranges (SFC_Tree::distTreePartition, src/tsort.cpp:229-508), node
 * ownership, and the ghost exchange readFromGhostBegin/End + writeToGhostsBegin/End
 * (include/oda.tcc:212-435) as ncclSend/ncclRecv groups.  Every rank passes the SAME full tree;
 * rank r keeps a contiguous range of the tree order, cut so that the per-rank work (elements weighted by their
 * interpolation cost, hanging > regular) is balanced -- not the equal-count split.  nccl_id: 128 bytes from
 * dkt_nccl_unique_id() on rank 0, broadcast by the caller (MPI, torch.distributed, a file ...).
 * Vectors passed to dkt_matvec then hold this rank's OWNED nodes (dkt_sizes.n_nodes of them);
 * dkt_da_export_owned_ids gives their index in the single-rank DA order. */
int dkt_nccl_unique_id(void *out128);
int dkt_da_create_dist(int dim, int order, int max_depth, int sfc_mode, const uint32_t *elem_xyz, const uint8_t *elem_lev,
                       uint64_t n_elem, const double *ip0, const double *ip1, unsigned flags, int rank, int nranks,
                       const void *nccl_id, dkt_da **out);
/* Test hook of the peer-memory exchange (DKT_DIST_P2P): das[0..n-1] are the DKT_DIST_DRYRUN DAs of
 * ranks 0..n-1 of ONE partition, all created in this process; their exchange buffers are wired to each other directly
 * (no IPC, no NCCL), after which dkt_matvec works on them with device vectors - enqueue the matvec of every rank before
 * synchronising any of them.  No reference counterpart. */
int dkt_p2p_attach_local(dkt_da **das, int n);
int dkt_da_export_owned_ids(const dkt_da *da, uint32_t *ids);
/* send_counts[p] = owned nodes rank p ghosts, recv_counts[p] = this rank's ghosts owned by p (n_ranks each) */
int dkt_da_export_exchange(const dkt_da *da, uint64_t *send_counts, uint64_t *recv_counts);

int dkt_da_sizes(const dkt_da *da, dkt_sizes *out);
/* tree in DA order: what DA::getTreePartFront()/Back() bracket (include/oda.h:255-258) */
int dkt_da_export_elements(const dkt_da *da, uint32_t *xyz, uint8_t *lev);
/* DA::getTNCoords() (include/oda.h:252): node coordinates and levels in DA order; partitioned DA: of the local vector
 * [owned | ghosts], n_nodes + n_ghost_nodes entries */
int dkt_da_export_nodes(const dkt_da *da, uint32_t *xyz, uint8_t *lev);
/* DA::getBoundaryNodeIndices() (include/oda.h:264); partitioned DA: the owned boundary nodes, local indices
 * (dkt_sizes.n_boundary of them) */
int dkt_da_export_boundary(const dkt_da *da, uint32_t *ids);
/* flat tables (no reference counterpart - the reference re-discovers them every matvec,
 * FEM/include/matvec.h:244-548).  Visited elements are stored regular-first: entries
 * [0, n_mv_elem - n_hanging) have every lattice node, the last n_hanging have hanging nodes.
 * mv_xyz[n_mv_elem*dim], mv_lev[n_mv_elem], e2n[n_mv_elem*N] (0xFFFFFFFF = hanging),
 * pnode[n_hanging*N] (0xFFFFFFFF = absent or level != L-1), child[n_hanging].
 * Any pointer may be NULL. */
int dkt_da_export_tables(const dkt_da *da, uint32_t *mv_xyz, uint8_t *mv_lev, uint32_t *e2n, uint32_t *pnode,
                         uint8_t *child);

/* Replaces feMatrix<LeafT,dim>::matVec(const VECType*in, VECType*out, double scale)
 * (FEM/include/feMatrix.h:190-259) including fem::matvec (FEM/include/matvec.h:232) for one
 * rank.  in/out have n_nodes doubles in DA order.  Device vectors: the kernels are enqueued on the DA's
 * stream (dkt_da_stream / dkt_da_set_stream) and the call returns without waiting; the caller orders that
 * stream after the producer of `in` and before the consumer of `out`. */
int dkt_matvec(dkt_da *da, const dkt_op *op, const double *in, double *out, double scale, unsigned flags);
/* dof > 1 (include/oda.h:296-322: vectors interleaved per node, [abc][abc]..; FEM/include/feMatrix.h m_uiDof): in/out hold
 * dof * n_nodes doubles, the elemental operator acts on every component.  The reference's own traversal stops at dof == 1
 * ("matvec only supports dof==1 right now", FEM/include/matvec.h:27), so this is checked against dof scalar matvecs. */
int dkt_matvec_dof(dkt_da *da, const dkt_op *op, const double *in, double *out, double scale, unsigned flags, int dof);

/* Conjugate gradients with every vector resident in HBM: the solver of the reference's example
 * operator, HeatMat::cgSolve(x, b, max_iter, tol) (FEM/examples/src/heatMat.cpp:165-325) - same
 * recurrences, same stopping rule |r|_inf / |b|_inf <= tol.  x (initial guess in, solution out) and b
 * are host or device vectors as selected by `flags` (DKT_VEC_*); *tol returns the achieved residual,
 * *iters the iterations done, *status 0 if converged within max_iter, 1 otherwise (the reference's
 * return value).  Each iteration is one dkt_matvec plus two fused vector kernels. */
int dkt_cg_solve(dkt_da *da, const dkt_op *op, double *x, const double *b, int max_iter, double *tol, double scale,
                 unsigned flags, int *iters, int *status);

/* Device time of the kernels of the most recent dkt_matvec on this DA, in milliseconds
 * (CUDA events on the DA's stream; excludes H2D/D2H). */
int dkt_last_kernel_ms(dkt_da *da, float *ms);
/* The CUDA stream (cudaStream_t) the DA launches on, for callers that time with events. */
void *dkt_da_stream(dkt_da *da);
/* Launch on the caller's stream from now on (e.g. torch's current stream); NULL restores the
 * DA's own stream.  The caller keeps the stream alive. */
int dkt_da_set_stream(dkt_da *da, void *cuda_stream);
/* The ghost exchanges of ot::DA on their own - readFromGhostBegin/End, writeToGhostsBegin/End (reference include/oda.h:300-322,
 * include/oda.tcc:212-435) - on a DEVICE vector of n_nodes + n_ghost_nodes doubles in the ghosted layout [owned | ghosts].
 * begin queues the NCCL exchange on the DA's exchange stream behind the work already queued on its stream, end makes the
 * stream wait for it (asynchronous with respect to the host).  read: owners -> ghost copies.  write: ghost entries -> their
 * owners, ADDED to the owned entries.  No-ops on an unpartitioned DA, like the reference at one rank.  dkt_matvec performs
 * both exchanges itself; these entry points serve callers that assemble or post-process ghosted vectors on their own. */
int dkt_ghost_read_begin(dkt_da *da, double *vec);
int dkt_ghost_read_end(dkt_da *da, double *vec);
int dkt_ghost_write_begin(dkt_da *da, double *vec);
int dkt_ghost_write_end(dkt_da *da, double *vec);
/* The same on HOST vectors of n_nodes + n_ghost_nodes doubles (blocking): what an application hands to
 * DA::readFromGhostBegin/End and writeToGhostsBegin/End.  read: fills the ghost segment; write: adds the ghost segment's
 * values to their owners' entries in the owned segment. */
int dkt_ghost_read_host(dkt_da *da, double *vec);
int dkt_ghost_write_host(dkt_da *da, double *vec);
/* Diagnostics of the chunked tables: out[0..4] = regular per-element sets {chunks, units per chunk, max nodes per
 * chunk, max run length, total chunk nodes}, out[5..9] = the same for the hanging per-element sets, out[10..14] for
 * the sibling-family sets (unit = family), out[15] = elements inside sibling families. */
int dkt_da_chunk_info(const dkt_da *da, uint64_t out[16]);
/* Number of kernels launched by this library in the calling process since load. */
uint64_t dkt_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
