/* ralf_b200 C ABI -- the drop-in boundary of the B200-native RALF hot path.
 *
 * The reference (CyberAgentAILab/RALF) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md 8b); each entry point below names the reference code whose arithmetic it replaces.
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (in practice the torch caching allocator), including workspaces;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *   - return 0 on success or a negative RalfStatus; no C++ exception crosses the ABI;
 *     the last CUDA error string is available from ralf_last_cuda_error();
 *   - sm_100a only: there is no CPU or other-arch fallback behind these symbols.
 *
 * "split" tensors: a bf16 tensor stored as two planes, hi = bf16(x) and lo = bf16(x - hi); the
 * lo plane starts `*_plane` ELEMENTS after the hi plane.  GEMMs consume split operands and (with
 * npass = 3: hi*lo + lo*hi + hi*hi, fp32 accumulate) reproduce an fp32 product to ~2^-17.
 */
#ifndef RALF_B200_H_
#define RALF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RalfStatus {
  RALF_OK = 0,
  RALF_ERR_SHAPE = -1,  /* unsupported / inconsistent shape argument            */
  RALF_ERR_ALIGN = -2,  /* pointer or stride not aligned as required            */
  RALF_ERR_NULL = -3,   /* required pointer is NULL                             */
  RALF_ERR_CUDA = -4,   /* CUDA runtime error (see ralf_last_cuda_error)        */
  RALF_ERR_DRIVER = -5, /* driver entry point / tensor-map encoding failed      */
  RALF_ERR_ARCH = -6,   /* device is not compute capability 10.x                */
  RALF_ERR_WORKSPACE = -7 /* workspace too small                                */
} RalfStatus;

const char* ralf_last_cuda_error(void);
/* 0 if device `dev` is sm_100 (B200), RALF_ERR_ARCH otherwise. */
int ralf_check_device(int dev);
int ralf_version(void);

/* ---------------------------------------------------------------------------------------------
 * K1. Exact maximum-inner-product top-k (replaces faiss.IndexFlat(d, METRIC_INNER_PRODUCT).search
 * reached from image2layout/train/models/retrieval/retriever.py:79-84,193-213).
 *
 * gallery [n, d] fp32 row-major (as the reference hands it to FAISS: NOT normalised), queries
 * [q, d] fp32.  Returns for every query the k largest inner products in descending order, ties
 * broken towards the lower index (FAISS leaves ties unspecified).  Scores are the canonical fp32
 * dot product defined in oracle/knn_oracle.c (lane-strided fmaf partials + xor butterfly), so
 * indices AND scores are bit-reproducible on the CPU.  out_idx is int64 like FAISS labels, with
 * `index_base` added (gallery shard offset for the multi-GPU path); missing results (n < k) are
 * idx = -1, score = -inf.
 *
 * Phase 1 streams the gallery once through a TF32 tcgen05 GEMM fused with a per-query running
 * top-C candidate filter (C = 32 for k <= 24, 64 for k <= 48); phase 2 merges candidates and
 * re-scores them exactly.  certified[q] (optional) is set to 1 when the TF32 error bound
 * 2^-9 * |query| * gallery_max_norm proves no true top-k item can have been missed by phase 1.
 * ------------------------------------------------------------------------------------------- */
size_t ralf_knn_workspace_bytes(int n, int d, int q, int k);
int ralf_knn_topk(const float* gallery, int n, int d, const float* queries, int q, int k,
                  long long index_base, float gallery_max_norm, long long* out_idx,
                  float* out_score, int* certified, void* workspace, size_t workspace_bytes,
                  void* stream);
/* Exact CUDA-core scan with the same canonical score (no tensor cores): the certified fallback
 * for queries phase 1 cannot certify, and an independent check of ralf_knn_topk. */
int ralf_knn_topk_exact(const float* gallery, int n, int d, const float* queries, int q, int k,
                        long long index_base, long long* out_idx, float* out_score,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Merge `parts` per-shard result lists ([parts, q, k] scores / indices, e.g. after an NCCL
 * all-gather) into the global top-k with the same ordering rule. */
int ralf_knn_merge(const float* part_score, const long long* part_idx, int parts, int q, int k,
                   long long* out_idx, float* out_score, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2. Dense contraction  D[M,N] = A[M,K] . W[N,K]^T  with fused epilogue (replaces every
 * nn.Linear / nn.Conv2d(+BatchNorm, folded) on the path: common/common.py:26-41,
 * common/attention.py:20-26,46-47, common/image.py:39-83, fid/model.py:66-68).
 * ------------------------------------------------------------------------------------------- */
typedef struct RalfGemmArgs {
  const void* A;      /* split bf16 [planes, M, K], K contiguous, row stride lda          */
  long long a_plane;  /* elements between hi and lo plane of A                            */
  int lda;
  const void* W;      /* split bf16 [planes, N, K] (nn.Linear weight layout)              */
  long long w_plane;
  int ldw;
  int M, N, K;
  int npass;          /* 3 = bf16x3 (fp32-faithful), 1 = plain bf16                       */
  int block_n;        /* 0 = auto, else 64 / 128 / 256                                    */
  const float* bias;  /* [N] or NULL                                                      */
  int act;            /* 0 none, 1 ReLU, 2 GELU(erf); applied after bias, before residual */
  int post_relu;      /* ReLU after the residual add (ResNet bottleneck tail)             */
  const float* res;   /* fp32 residual [*, res_ld] or NULL                                */
  const void* res_split; /* split bf16 residual or NULL                                   */
  long long res_plane;
  int res_ld;
  int res_row_mod;    /* >0: residual row = row % res_row_mod (broadcast table, e.g. PE)  */
  float* out_f32;     /* fp32 output or NULL                                              */
  void* out_split;    /* split bf16 output or NULL                                        */
  long long out_plane;
  int out_split_lo;   /* also write the lo plane                                          */
  int out_ld;         /* row stride (elements) of both outputs                            */
  int out_col0;       /* first output column                                              */
  int rows_per_group; /* >0: out_row = (r / rpg) * group_stride + group_offset + r % rpg  */
  int group_stride;
  int group_offset;
} RalfGemmArgs;
int ralf_gemm(const RalfGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RALF_B200_H_ */
