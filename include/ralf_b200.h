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
/* Device-side exactness guarantee (no host round trip; safe inside a CUDA graph): every query with certified[q] == 0
 * after ralf_knn_topk is re-run through the exact scan and its rows of out_idx / out_score are overwritten;
 * certified[q] becomes 2.  With nothing to fix the three launches return after one load each.  Replaces the
 * "IndexFlat is exact" property of the reference's FAISS index (retrieval/retriever.py:79-84,200-202) for inputs the
 * TF32 bound cannot certify (e.g. more than C near-duplicate rows around the cut).  Same workspace as ralf_knn_topk. */
int ralf_knn_fixup_exact(const float* gallery, int n, int d, const float* queries, int q, int k,
                         long long index_base, int* certified, long long* out_idx, float* out_score,
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
  void* out_kv24;     /* decoder cross-attention K/V cache rows in the 24-bit format, or NULL (N must be 512):
                       * row r = 1536 bytes [K hi 256 x u16 | V hi 256 x u16 | K lo 256 x u8 | V lo 256 x u8]; a value
                       * is the fp32 result rounded to 24 bits, hi = bits 31..16, lo = bits 15..8 (16-bit mantissa,
                       * the accuracy class of the bf16x3 product that made it) -- 3 instead of 4 bytes per element on
                       * the stream the decode loop is bound by */
  int out_kv_fmt;     /* 0 / 24: the 24-bit format above.  16: rows of 1088 bytes
                       * [K 256 x u16 | V 256 x u16 | 8 x (K scale, V scale) f32]: per (row, head) the 32 values are
                       * stored as offset-binary 16-bit integers q + 32768, q = rint(x * 32767 / amax), next to the scale
                       * amax / 32767 -- 2.125 bytes per element; error 2^-16 of the head's largest value (measured on the
                       * reference goldens: step logits within 8e-5 of scale, token ids identical;
                       * profiles/r2_precision_study.json) */
  void* splitk_ws;    /* optional workspace: lets a plain fp32-output GEMM with few M x N tiles and a long K (weight
                       * gradients: K = all rows of the batch) run split-K (k slices -> partials -> deterministic sum) */
  size_t splitk_ws_bytes; /* >= ralf_gemm_splitk_workspace_bytes(M, N, K) */
} RalfGemmArgs;
size_t ralf_gemm_splitk_workspace_bytes(int M, int N, int K);
int ralf_gemm(const RalfGemmArgs* args, void* stream);
/* Stride-1 "same" convolution (nn.Conv2d(C, N, KH, padding=KH/2), BatchNorm folded; common/image.py:39-83 -- the 3x3
 * convolutions of the ResNet50 bottlenecks and of the FPN) as an IMPLICIT GEMM: args->A is the NHWC split activation
 * [planes][B*H*W, C] itself, each k-block of the main loop is one (filter tap, 64-channel slice) fetched by a 5-D TMA
 * box whose out-of-image part is zero-filled (= the padding); no im2col buffer exists.
 * args->W = [planes][N, KH*KW*C] with k = (kh*KW + kw)*C + c; args->M = B*H*W; args->K = KH*KW*C; C % 64 == 0. */
/* Decode-loop residual GEMM with the FOLLOWING LayerNorm fused in (nn.TransformerDecoderLayer, norm_first:
 * x = x + sublayer(..); h = norm(x), models/common/common.py:84-135): x_new = A . W^T + bias + res -> out_f32 (may alias
 * res), LayerNorm(x_new; gamma, beta, eps) -> ln_split as split bf16 rows [2][M, 256].  N must be 256, K % 64 == 0,
 * npass 3; args supplies A, W, bias, res / res_ld, out_f32 / out_ld, every other epilogue field unset.  The eight
 * 32-column tiles of a row block run as one thread-block cluster and exchange row statistics through distributed shared
 * memory. */
int ralf_gemm_res_ln(const RalfGemmArgs* args, const float* gamma, const float* beta, float eps, void* ln_split,
                     long long ln_plane, void* stream);
int ralf_conv_gemm(const RalfGemmArgs* args, int B, int H, int W, int C, int KH, int KW, void* stream);
/* The same with stride 1 or 2 (ResNet's stride-2 3x3 and 1x1 downsample convolutions, torchvision resnet50 via
 * models/common/image.py:90-120): args->M = B*Ho*Wo, Ho = (H + 2*(KH/2) - KH) / stride + 1; a tap's TMA box skips every
 * other input position through the tensor map's element strides, so no im2col rows are materialised. */
int ralf_conv_gemm_strided(const RalfGemmArgs* args, int B, int H, int W, int C, int KH, int KW, int stride, void* stream);
/* ResNet50 stem (7x7 / stride 2 / pad 3 on the 4-channel canvas, common/image.py:69-77) in space-to-depth form:
 * ralf_stem_s2d turns the fp32 NCHW image [B,4,H,W] (H, W even) into a zero-bordered NHWC split buffer
 * [planes][B, H/2+3, W/2+3, 16] (channel = (dy*2+dx)*4 + c); ralf_stem_gemm runs the equivalent 4x4 / stride 1
 * convolution over it as an implicit GEMM: args->A = that buffer, args->W = [planes][N, 256] with
 * k = kh'*64 + kw'*16 + (dy*2+dx)*4 + c, args->M = B*Ho*Wo (Ho = H/2, Wo = W/2 <= 128), args->K = 256. */
int ralf_stem_s2d(const float* img, int B, int H, int W, void* out, long long out_plane, void* stream);
int ralf_stem_gemm(const RalfGemmArgs* args, int B, int Ho, int Wo, void* stream);
/* D = LayerNorm(x)[M,256] . W^T with the LayerNorm computed inside the GEMM (decode path: K = 256,
 * npass = 3; every (n-tile, 128-row tile) CTA normalises its rows).  x fp32 [M, 256] (row stride ldx); `args` supplies W and the epilogue (its A fields are ignored).
 * Replaces nn.LayerNorm + nn.Linear pairs of the pre-LN decoder layer / LM head (common/common.py:26-41). */
int ralf_gemm_ln(const float* x, int ldx, const float* gamma, const float* beta, float eps,
                 const RalfGemmArgs* args, void* stream);

/* Fused decode-step chain (csrc/decode_chain.cu): up to 4 row-local stages
 *     [LayerNorm | split rows from global | previous stage's output]  ->  Linear (W split bf16 [2, n_out, k_in], bf16x3)
 *     -> + bias -> ReLU -> + residual row -> {new residual row, global fp32, next stage's operand}
 * for the single new token of each of B canvases, in ONE kernel: a CTA owns 16 canvases, streams the chain's weights
 * through a TMA ring and keeps the residual row / FFN hidden layer in shared memory.  Replaces the per-op launches of
 * one pre-LN decoder layer step (common/common.py:26-41,84-135; retrieval_augmented_autoreg.py:271-297):
 *   LN1 -> in_proj                                   (1 stage)
 *   out_proj + x -> LN2 -> cross-attention query     (2 stages)
 *   out_proj + x -> LN3 -> linear1 -> ReLU -> linear2 + x  [-> next layer's LN1 -> in_proj | final LN -> LM head]
 * x: residual rows fp32 [B, 256] (row stride ldx), read once at the start (needed by LayerNorm / add_x stages).
 * k_in is 256 (LayerNorm / split-row inputs) or 1024 (operand = previous stage's n_out = 1024 with out_operand). */
typedef struct RalfChainStage {
  const void* W;      /* split bf16 [2, n_out, k_in], K contiguous, row stride ldw */
  long long w_plane;
  int ldw;
  int n_out;
  int k_in;
  int in_mode;        /* 0: previous stage's out_operand, 1: LayerNorm(residual row; gamma, beta, eps), 2: in_split rows */
  const float* gamma;
  const float* beta;
  float eps;
  const void* in_split; /* split bf16 [2, B, 256] (e.g. an attention output) */
  long long in_plane;
  int in_ld;
  const float* bias;  /* [n_out] or null */
  int act;            /* 0 none, 1 ReLU */
  int add_x;          /* + residual row (n_out == 256) */
  int to_x;           /* the result becomes the residual row (n_out == 256) */
  int out_operand;    /* the result (split) is the next stage's operand; n_out == 1024 */
  float* out_f32;     /* global fp32 [B, out_ld] or null */
  int out_ld;
} RalfChainStage;
int ralf_decode_chain(const float* x, int ldx, int B, const RalfChainStage* stages, int n_stages, void* stream);


/* ---------------------------------------------------------------------------------------------
 * K3. Non-GEMM ops of the forward / generate path (fp32 CUDA-core kernels, csrc/nn_kernels.cu).
 * "split" outputs feed the next GEMM directly.
 * ------------------------------------------------------------------------------------------- */
/* nn.LayerNorm(D) over rows of x (row stride in_ld); D % 32 == 0, D <= 1024. */
int ralf_layernorm(const float* x, long long in_ld, const float* gamma, const float* beta, float eps,
                   int M, int D, float* out_f32, void* out_split, long long out_plane, void* stream);
/* softmax(q k^T * scale + masks) v per (batch, head); replaces nn.MultiheadAttention's core inside
 * nn.TransformerEncoderLayer / DecoderLayer (common/common.py:26-35) and common/attention.py:64-69.
 * q row (b,t) at q + (b*Tq+t)*ldq + h*head_dim; k/v row (b,j) at k + (b*Tk+j)*ldk + h*head_dim.
 * key_padding_mask: uint8 [B,Tk], nonzero = padded key; causal: keys j > t masked. */
int ralf_attention(const float* q, int ldq, const float* k, const float* v, int ldk,
                   const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                   int causal, float scale, void* out_split, long long out_plane, float* out_f32, int ldo,
                   void* stream);
/* One query per (batch, head) against a K/V cache: the KV-cached replacement of the reference's
 * full-prefix recompute (retrieval_augmented_autoreg.py:271-297).  K/V row j of batch b at
 * base + (b*kv_bstride + j)*ldk + h*head_dim. */
int ralf_attention_decode(const float* q, int ldq, const float* k, const float* v, long long kv_bstride,
                          int ldk, const unsigned char* key_padding_mask, int mask_ld, int Tk, int B, int H,
                          int head_dim, float scale, void* out_split, long long out_plane, int ldo,
                          void* stream);
/* ralf_attention_decode over a 24-bit K/V cache written by ralf_gemm(out_kv24): rows (b*kv_bstride + j) of 1536 bytes,
 * 8 heads x 32, no mask (memory cross-attention of the decode loop). */
int ralf_attention_decode_kv24(const float* q, int ldq, const void* kv24, long long kv_bstride, int Tk, int B, int H,
                               float scale, void* out_split, long long out_plane, int ldo, void* stream);
/* The same over the 16-bit per-head-scaled cache (ralf_gemm out_kv24 with out_kv_fmt = 16, rows of 1088 bytes). */
int ralf_attention_decode_kv16(const float* q, int ldq, const void* kv16, long long kv_bstride, int Tk, int B, int H,
                               float scale, void* out_split, long long out_plane, int ldo, void* stream);
/* Self-attention decode step with the cache append fused in: qkv [B, 3*H*head_dim] (this step's fused
 * projection), K/V caches [B, S, H*head_dim]; writes row `pos` of both caches, then attends over keys 0..pos. */
int ralf_attention_decode_append(const float* qkv, int ldqkv, float* kcache, float* vcache, int S, int pos,
                                 const unsigned char* key_padding_mask, int mask_ld, int B, int H, int head_dim,
                                 float scale, void* out_split, long long out_plane, int ldo, void* stream);
/* ResNet50 stem (common/image.py:69-77): image fp32 NCHW [B,4,H,W] -> split im2col rows
 * [B*Ho*Wo, KP] for the 7x7/s2/p3 conv, k = (kh*7+kw)*4 + c, KP = 200 (zero padded). */
int ralf_stem_im2col(const float* img, int B, int H, int W, int KP, void* out, long long out_plane,
                     void* stream);
/* NHWC split im2col: [B,H,W,C] -> [B*Ho*Wo, KH*KW*C], k = (kh*KW+kw)*C + c; C % 8 == 0. */
int ralf_im2col(const void* in, long long in_plane, int B, int H, int W, int C, int KH, int KW, int stride,
                int pad, void* out, long long out_plane, void* stream);
int ralf_maxpool3x3s2(const void* in, long long in_plane, int B, int H, int W, int C, void* out,
                      long long out_plane, void* stream);
/* FPN merge (common/image.py:103-111): fused[:, 0:C] = nearest-upsampled c5; sum = up + c4. */
int ralf_fpn_merge(const float* c5, const float* c4, int B, int h5, int w5, int h4, int w4, int C,
                   void* fused, long long fused_plane, int ldf, void* sum, long long sum_plane, void* stream);
/* out[map(r), :] = in[r, :] * scale + add + table[r % tab_mod, :]  (in/table may be NULL). */
int ralf_rows_affine(const float* in, long long in_ld, int M, int D, float scale, float add,
                     const float* table, int tab_mod, int rows_per_group, int group_stride, int group_offset,
                     float* out_f32, void* out_split, long long out_plane, int out_ld, void* stream);
/* out[b*S+s, :] = emb[tok[b*tok_ld + tok_col + s], :] * scale + pe[pos0 + s, :]
 * (common/common.py:99-100, positional_encoding.py:94-107). */
int ralf_embed(const long long* tok, long long tok_ld, int tok_col, int B, int S, const float* emb, int D,
               float scale, const float* pe, int pos0, float* out, void* stream);
/* FIDNetV3 input rows cat[fc_bbox(cx,cy,w,h), emb_label[label]] (fid/model.py:95-101). */
int ralf_fid_embed(const float* cx, const float* cy, const float* w, const float* h, const long long* label,
                   int rows, int D, const float* fc_w, const float* fc_b, const float* emb, void* out,
                   long long out_plane, void* stream);
/* Exemplar fetch as an index gather (replaces RetrievalDatasetWrapper.__getitem__'s 16 row reads,
 * helpers/retrieval_dataset_wrapper.py:89-148): packed table [n_table, row_elems] fp32, row_elems = 6*E
 * (label, mask, center_x, center_y, width, height), idx [rows] global ids (table holds ids
 * [index_base, index_base + n_table)); missing (-1) -> empty layout. */
int ralf_gather_layouts(const float* table, const long long* idx, int rows, int row_elems, long long n_table,
                        long long index_base, float* out, void* stream);
/* FIDNetV3 input rows + key-padding mask [nseq, E+1] straight from packed layouts [nseq, 6, E]. */
int ralf_fid_embed_packed(const float* packed, int nseq, int E, int D, const float* fc_w, const float* fc_b,
                          const float* emb, int num_labels, void* out, long long out_plane,
                          unsigned char* pad_mask, void* stream);
/* Greedy step tail (retrieval_augmented_autoreg.py:281-297, helpers/sampling.py:24-25): mask the
 * vocabulary, argmax (first maximum), append to seq[:, pos], update the pad mask, embed the token. */
int ralf_argmax_next(const float* logits, int ldl, int B, int V, const unsigned char* allowed, long long* seq,
                     int seq_ld, int pos, unsigned char* pad_mask, int mask_ld, long long pad_id,
                     const float* emb, int D, float scale, const float* pe, float* x_next, void* stream);
/* Constrained / stochastic step tail: replaces the per-sample python loops of
 * DECODE_SPACE_RESTRICTION (layoutformerpp/decoding_space_restriction.py:5-106) and helpers/sampling.py:18-68.
 * forced: int32 [B, forced_ld] or NULL; forced[b, step] >= 0 -> emit that token.  mode: 0 greedy, 1 random, 2 top_k,
 * 3 top_p, 4 gumbel.  uniform: fp32 [B] in [0,1) (modes 1-4), noise: fp32 [B, noise_ld] uniforms (mode 4).  The draw is
 * the inverse CDF over the kept tokens in (logit desc, index asc) order.  V <= 1024.  Tail as ralf_argmax_next. */
int ralf_sample_next(const float* logits, int ldl, int B, int V, const unsigned char* allowed, const int* forced,
                     int forced_ld, int step, int mode, float temperature, int top_k, float top_p,
                     const float* uniform, const float* noise, int noise_ld, long long* seq, int seq_ld, int pos,
                     unsigned char* pad_mask, int mask_ld, long long pad_id, const float* emb, int D, float scale,
                     const float* pe, float* x_next, void* stream);
/* nn.CrossEntropyLoss(label_smoothing=eps, ignore_index) with mean reduction over logits [M, V]
 * (retrieval_augmented_autoreg.py:140-142,213-214); workspace = 2*M floats; out_loss = 1 float. */
int ralf_ce_label_smooth(const float* logits, int ldl, const long long* targets, int M, int V, float eps,
                         long long ignore_index, float* workspace, float* out_loss, void* stream);
/* Append this step's K,V (columns [D,3D) of the fused QKV row) to the self-attention cache [B,S,D]. */
int ralf_kv_append(const float* qkv, int B, int D, float* kcache, float* vcache, int S, int pos, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4. Training step (train.py:431-459): backward of the ops above, BatchNorm in training mode, and the
 * optimizer.  The backward contractions dX = dY.W and dW = dY^T.X reuse ralf_gemm on K-major operands
 * produced by ralf_to_split / ralf_transpose_to_split.  csrc/train_kernels.cu, csrc/conv_train_kernels.cu.
 * ------------------------------------------------------------------------------------------- */
/* [R, C] fp32 or split -> split [C, R] with row pitch ld_out (>= R, multiple of 8). */
int ralf_transpose_to_split(const float* in_f32, const void* in_split, long long in_plane, long long ld_in, int R,
                            int C, void* out, long long out_plane, long long ld_out, void* stream);
int ralf_to_split(const float* in, long long total, void* out, long long out_plane, void* stream);
/* fp32 [R, C] -> split [2, R, out_ld] and its transpose split [2, C, outT_ld] in one pass (dY for dgrad and wgrad). */
int ralf_split_and_transpose(const float* in, long long ld_in, int R, int C, void* out, long long out_plane, long long out_ld,
                             void* outT, long long outT_plane, long long outT_ld, void* stream);
/* Multi-tensor weight refresh after an optimiser step: every task turns an fp32 weight [R, C] into the split GEMM operand
 * W [2, R, w_ld] and its transpose W^T [2, C, wt_ld] (dgrad operand).  `tasks_dev` is a DEVICE array sorted by tile0
 * (tile0 = running sum of ceil(R/32)*ceil(C/32)); one launch of total_tiles CTAs. */
typedef struct RalfRefreshTask {
  const float* src;
  long long ld_in;
  int R, C;
  void* w;
  long long w_plane, w_ld;
  void* wt;
  long long wt_plane, wt_ld;
  int tile0;
  int tiles_x;
} RalfRefreshTask;
int ralf_refresh_operands(const RalfRefreshTask* tasks_dev, int ntasks, int total_tiles, void* stream);
/* out[c] (+)= sum_r in[r, c]  (bias gradients; deterministic order). */
/* workspace: ralf_colsum_workspace_bytes(M, C) bytes (0 = small M: single pass) or null; both paths are deterministic. */
size_t ralf_colsum_workspace_bytes(int M, int C);
int ralf_colsum(const float* in, long long ld, int M, int C, float* out, int accumulate, float* workspace, void* stream);
/* nn.LayerNorm backward; dx = add_to + dLN; workspace = 2*D*min(ceil(M/8), 4*SMs) floats. */
int ralf_layernorm_bwd(const float* x, long long x_ld, const float* dy, const float* gamma, float eps, int M, int D,
                       const float* add_to, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream);
/* ---- dropout (training; nn.Dropout / nn.MultiheadAttention(dropout=p) of the reference's encoder / decoder layers and
 * positional encodings, retrieval_augmented_autoreg.py:105,116-126; common/common.py:26-35,216; positional_encoding.py:67-107).
 * Masks are counter-based: element idx of call site `site` is kept iff SplitMix64(*seed ^ f(site), idx) >> 40 >= p * 2^24;
 * `seed` is a DEVICE pointer (one value per step) so captured graphs draw fresh masks per replay.  Forward and backward
 * recompute the mask; nothing is stored. */
/* y = (res ? res : 0) + dropout(x): x fp32 [total] or split (in_plane); y fp32 and / or split.  In place allowed. */
int ralf_dropout(const float* in_f32, const void* in_split, long long in_plane, const float* res, long long total,
                 const unsigned long long* seed, unsigned int site, float p, float* out_f32, void* out_split,
                 long long out_plane, void* stream);
/* Keep mask (1 = kept) of elements [0, total) of a site: test / debugging aid. */
int ralf_dropout_mask(const unsigned long long* seed, unsigned int site, float p, long long total, unsigned char* out,
                      void* stream);
/* ralf_attention with dropout on the attention probabilities: O = (softmax(S) o M / (1-p)) V; mask element index
 * ((b*H + h)*Tq + t)*Tk + j.  Always the CUDA-core kernel.  lse_out (optional, p > 0 only): B*H*Tq floats, the row
 * log-sum-exp, which ralf_attention_bwd_dropout(lse_given = 1) then reads from lse_ws instead of recomputing it. */
int ralf_attention_dropout(const float* q, int ldq, const float* k, const float* v, int ldk,
                           const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim, int causal,
                           float scale, void* out_split, long long out_plane, float* out_f32, int ldo,
                           const unsigned long long* seed, unsigned int site, float p, float* lse_out, void* stream);
int ralf_attention_bwd_dropout(const float* q, int ldq, const float* k, const float* v, int ldk,
                               const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                               int causal, float scale, const void* o_split, long long o_plane, const float* dO, int ldo,
                               float* lse_ws, float* delta_ws, float* dq, int lddq, float* dk, float* dv, int lddk,
                               const unsigned long long* seed, unsigned int site, float p, int lse_given, void* stream);
/* Backward of ralf_attention (same addressing); lse_ws / delta_ws: B*H*Tq floats each. */
int ralf_attention_bwd(const float* q, int ldq, const float* k, const float* v, int ldk,
                       const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim, int causal,
                       float scale, const void* o_split, long long o_plane, const float* dO, int ldo, float* lse_ws,
                       float* delta_ws, float* dq, int lddq, float* dk, float* dv, int lddk, void* stream);
/* dlogits of ralf_ce_label_smooth (fwd_workspace = the forward's workspace), scaled by `scale`. */
int ralf_ce_label_smooth_bwd(const float* logits, int ldl, const long long* targets, int M, int V, float eps,
                             long long ignore_index, const float* fwd_workspace, float scale, float* dlogits, int ldd,
                             void* stream);
int ralf_relu_bwd(float* dy, const void* y_split_hi, long long total, void* stream);
int ralf_gelu_fwd(const float* z, long long total, void* out_split, long long out_plane, void* stream);
int ralf_gelu_bwd(float* dy, const float* z, long long total, void* stream);
int ralf_axpy(float* a, const float* b, float alpha, long long total, void* stream);
/* dst[r,:] (+)= scale * src[map(r),:] -- backward of ralf_rows_affine / concatenations. */
int ralf_rows_gather(const float* src, long long src_ld, int M, int D, float scale, int rows_per_group,
                     int group_stride, int group_offset, float* dst, int accumulate, void* stream);
/* V = rows of the embedding table.  Deterministic (gather per vocabulary row, no atomics). */
int ralf_embed_bwd(const long long* tok, long long tok_ld, int tok_col, int B, int S, const float* dy, int D,
                   float scale, float* demb, int V, void* stream);
/* torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW over flat buffers (train.py:450-454). */
int ralf_grad_norm(const float* grads, long long n, float* workspace /* 1024 floats */, float* out_norm, void* stream);
int ralf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                    const float* grad_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream);
/* Same, with the per-step scalars in device memory (dyn = {lr scale, 1 - beta1^t, 1 - beta2^t}) so that a captured
 * CUDA graph of the whole training step can be replayed. */
int ralf_adamw_step_dyn(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                        const float* grad_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                        float weight_decay, const float* dyn, void* stream);
/* BatchNorm2d in training mode on NHWC rows [M, C]: mode 0 = batch mean / rstd (+ running-stat update),
 * mode 1 = (sum a, sum a*xhat) for the backward; workspace = 2*C*ceil(M/256) floats. */
int ralf_bn_colstats(const float* a, const float* z, const float* mean, const float* rstd, int mode, int M, int C,
                     float eps, float momentum, float* out0, float* out1, float* running_mean, float* running_var,
                     float* workspace, void* stream);
int ralf_bn_apply(const float* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  const void* res_split, long long res_plane, int relu, int M, int C, void* out_split,
                  long long out_plane, float* out_f32, void* stream);
int ralf_bn_bwd_apply(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                      const float* sum_dy, const float* sum_dy_xhat, int M, int C, float* dz, void* stream);
/* Backward data movement of the convolutions / pooling / FPN upsample. */
int ralf_col2im(const float* dcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad, float* dx,
                int accumulate, void* stream);
/* Deterministic (first-maximum tap recorded per output in `workspace`, B*Ho*Wo*C bytes; gather per input pixel); dx is
 * written completely. */
int ralf_maxpool3x3s2_bwd(const void* x_split, long long x_plane, const float* dy, int B, int H, int W, int C,
                          float* dx, void* workspace, void* stream);
int ralf_upsample_nearest_bwd(const float* dbig, long long ld_big, int B, int h5, int w5, int h4, int w4, int C,
                              float* dsmall, void* stream);
/* nn.Conv2d weight [N, C, T] <-> GEMM layout [N, T*C] (split, plus transposed copy) and gradient back. */
int ralf_conv_weight_to_gemm(const float* w, int N, int C, int T, int Kp, void* out, long long out_plane, void* outT,
                             long long outT_plane, int Np, void* stream);
int ralf_conv_grad_from_gemm(const float* dwg, int N, int C, int T, int ldg, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RALF_B200_H_ */
