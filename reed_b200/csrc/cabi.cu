// C-ABI surface of libreed_sm100.so (declared in include/reed_b200.h).  Plain pointers and sizes only; every
// call enqueues on the caller's stream and returns 0, or non-zero with a message in reed_last_error().
#include <stdarg.h>
#include "common.cuh"

namespace reed {

thread_local char g_err[512] = {0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int gemm_simt(int act_dtype, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D,
              int64_t ldd, int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st);
int gemm_tcgen05(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                 int d_dtype, int M, int N, int K, const EpiParams& ep, cudaStream_t st);
int outer_wgrad(int dy_dtype, const void* dy, int64_t ld_dy, int x_dtype, const void* x, int64_t ld_x, float* D, int64_t ldd,
                int N, int K, int B, int accumulate, float* db, cudaStream_t st);
int gemm_tcgen05_grouped(int mode, const void* const* A, int64_t lda, const void* const* B, int64_t ldb, int groups, int per_group,
                         void* D, int64_t ldd, int M, int N, int K, const EpiParams& ep, cudaStream_t st);
bool gemm_tcgen05_supported(int64_t lda, int64_t ldb, int64_t ldd, const void* A, const void* B, int M, int N, int K);
void gemm_tcgen05_force_cta_group(int cg);
void gemm_tcgen05_force_bn(int bn);
void gemm_tcgen05_reserve_sms(int n);
int attn_simt_fwd(int act_dtype, const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st);
int attn_simt_bwd(int act_dtype, const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                  float* delta, int B, int T, int H, int hd, cudaStream_t st);
int attn_mma_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st);
int attn_mma_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B,
                 int T, int H, int hd, cudaStream_t st);
bool attn_mma_supported(int T, int hd);
int attn_fa_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int hd, cudaStream_t st);
bool attn_fa_supported(int T, int hd);
int attn_fa_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta, int B, int T,
                int H, int hd, cudaStream_t st);
bool attn_fa_bwd_supported(int T, int hd);

}  // namespace reed

using namespace reed;

extern "C" int reed_version(void) { return 102; }   // 0.1.2: row backward entries take ld_dmod; reed_gemm_grouped, reed_outer_wgrad

extern "C" const char* reed_last_error(void) { return g_err; }

// 0 when the current device is an sm_100 part; fills name (optional) with the device name
extern "C" int reed_device_check(char* name, int name_len) {
  int dev = 0;
  REED_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  REED_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && name_len > 0) snprintf(name, name_len, "%s", prop.name);
  REED_REQUIRE(prop.major == 10, "libreed_sm100 needs an sm_100 device, found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
  return 0;
}

// SMs the persistent tensor-core GEMM grids leave free from now on (host-side planner state; 0 = use every SM).
// The data-parallel trainer raises it around backward so the NCCL all-reduce kernels of finished gradient buckets
// find SMs without pushing part of a GEMM grid into a second wave (train.py:401: DDP's overlapped all-reduce).
// Running count of tcgen05 GEMM kernel launches made through reed_gemm / reed_gemm_wgrad_bias (host-side bookkeeping):
// bench.py checks that the launches it times for the roofline are ALL the tensor-core GEMM launches of a step.
static long long g_tcgen05_launches = 0;
extern "C" int reed_gemm_tcgen05_launches(long long* out) {
  REED_REQUIRE(out != nullptr, "gemm_tcgen05_launches: null output");
  *out = g_tcgen05_launches;
  return 0;
}

extern "C" int reed_gemm_reserve_sms(int n) {
  REED_REQUIRE(n >= 0 && n <= 64, "gemm_reserve_sms: %d out of range", n);
  gemm_tcgen05_reserve_sms(n);
  return 0;
}

// backend: 0 auto (tcgen05 for bf16 activations when the shape allows, else SIMT), 1 force SIMT, 2 require tcgen05,
// 3 / 4 require tcgen05 with cta_group::1 / cta_group::2 (test knob; the planner normally picks)
extern "C" int reed_gemm(int act_dtype, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                         int b_mn_major, void* D, int64_t ldd, int d_dtype, int M, int N, int K, int epilogue,
                         const void* bias, const void* aux, int64_t ld_aux, const void* gate, int64_t ld_gate,
                         int rows_per_group, void* out2, int64_t ld_out2, int accumulate, int backend, void* stream) {
  REED_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative dimension");
  if (M == 0 || N == 0) return 0;
  REED_REQUIRE(epilogue >= kEpiNone && epilogue <= kEpiDSilu, "gemm: unknown epilogue %d", epilogue);
  REED_REQUIRE(!(accumulate && (epilogue != kEpiNone || d_dtype != kF32)), "gemm: accumulate needs fp32 D and no epilogue");
  REED_REQUIRE(!(epilogue == kEpiGateRes && (d_dtype != kF32 || !aux || !gate || rows_per_group <= 0)),
               "gemm: gate+residual epilogue needs fp32 D, residual, gate and rows_per_group");
  REED_REQUIRE(!((epilogue == kEpiDGelu || epilogue == kEpiDSilu) && !aux), "gemm: act' epilogue needs the saved pre-activation");
  EpiParams ep;
  ep.kind = epilogue; ep.bias = (const float*)bias; ep.aux = aux; ep.ld_aux = ld_aux; ep.gate = (const float*)gate;
  ep.ld_gate = ld_gate; ep.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; ep.out2 = out2; ep.ld_out2 = ld_out2;
  ep.accumulate = accumulate;
  ep.bias_grad = nullptr; ep.n_store = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int force_bn = backend >> 3;       // test knob: bits 3+ of `backend` pin the tile width (1/2/3 = 128/192/256)
  backend &= 7;
  const bool tc_ok = act_dtype == kBF16 && gemm_tcgen05_supported(lda, ldb, ldd, A, B, M, N, K) &&
                     (ld_aux % 4 == 0) && (ld_out2 % 4 == 0) && (ld_gate % 4 == 0);
  if (backend >= 2) REED_REQUIRE(tc_ok, "gemm: tcgen05 path required but shape/dtype unsupported (M=%d N=%d K=%d)", M, N, K);
  // a weight gradient over a handful of rows (adaLN / embedder linears: the contraction is the batch, K <= 64) is an
  // outer-product stream that writes M*N floats - the SIMT kernel does it at HBM speed, a 256-wide MMA tile would not
  const bool tiny_wgrad = a_mn_major && b_mn_major && K <= 64 && d_dtype == kF32 && epilogue == kEpiNone && !bias &&
                          backend < 2 && N % 4 == 0;
  if (backend != 1 && tc_ok && !tiny_wgrad) {
    gemm_tcgen05_force_cta_group(backend == 3 ? 1 : (backend == 4 ? 2 : 0));
    gemm_tcgen05_force_bn(force_bn == 1 ? 128 : (force_bn == 2 ? 192 : (force_bn == 3 ? 256 : 0)));
    ++g_tcgen05_launches;
    return gemm_tcgen05(A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, d_dtype, M, N, K, ep, st);
  }
  return gemm_simt(act_dtype, A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, d_dtype, M, N, K, ep, st);
}

// Weight and bias gradient of a Linear whose contraction is the batch (rows <= 64: the adaLN modulation / embedder linears):
// dW[n_out, k_in] (+)= dy^T x, db[n_out] += column sums of dy (db may be NULL).  dy [rows, n_out] fp32 or bf16, x [rows, k_in].
extern "C" int reed_outer_wgrad(const void* dy, int dy_dtype, int64_t ld_dy, const void* x, int x_dtype, int64_t ld_x, void* dW,
                                int64_t ldd, void* db, int n_out, int k_in, int rows, int accumulate, void* stream) {
  REED_REQUIRE(dy != nullptr && x != nullptr && dW != nullptr, "outer_wgrad: null operand");
  REED_REQUIRE(n_out > 0 && k_in > 0, "outer_wgrad: empty output");
  return outer_wgrad(dy_dtype, dy, ld_dy, x_dtype, x, ld_x, (float*)dW, ldd, n_out, k_in, rows, accumulate, (float*)db,
                     (cudaStream_t)stream);
}

// Several linears that share their input as ONE tensor-core launch (the adaLN-Zero modulation linears of all blocks,
// sit.py:125-133; bf16 operands, fp32 D, M <= 1024).  A / B: host arrays of `groups` device pointers.
//   mode 0: D[M, groups * per_group] = A[0] . [B_0; B_1; ...]^T + bias      (B_g: [per_group, K] row-major)
//   mode 1: D[M, N] (+)= sum_g A_g[M, per_group] . B_g[per_group, N]        (the input gradient of mode 0)
extern "C" int reed_gemm_grouped(int mode, const void* const* A, int64_t lda, const void* const* B, int64_t ldb, int groups,
                                 int per_group, void* D, int64_t ldd, int M, int N, int K, const void* bias, int accumulate,
                                 void* stream) {
  REED_REQUIRE(mode == 0 || mode == 1, "gemm_grouped: mode %d", mode);
  REED_REQUIRE(A != nullptr && B != nullptr && D != nullptr, "gemm_grouped: null operand table");
  REED_REQUIRE(M > 0 && N > 0 && K > 0 && groups > 0 && per_group > 0, "gemm_grouped: empty problem");
  EpiParams ep;
  ep.kind = kEpiNone; ep.bias = (const float*)bias; ep.aux = nullptr; ep.ld_aux = 0; ep.gate = nullptr; ep.ld_gate = 0;
  ep.rows_per_group = 1; ep.out2 = nullptr; ep.ld_out2 = 0; ep.accumulate = accumulate; ep.bias_grad = nullptr; ep.n_store = 0;
  ++g_tcgen05_launches;
  return gemm_tcgen05_grouped(mode, A, lda, B, ldb, groups, per_group, D, ldd, M, N, K, ep, (cudaStream_t)stream);
}

// qkv: [B, T, 3, H, hd] (act dtype); o: [B, T, H, hd]; lse: [B, H, T] fp32
extern "C" int reed_colsum(const void* src, int src_dtype, int64_t ld, void* out, int M, int N, void* stream);

// dW[n_out, k_in] (+)= dy^T x and db[n_out] += sum_tokens dy, in ONE weight-gradient GEMM: x_ext is the activation
// matrix [tokens, k_in + 8] whose column k_in holds ones (reed_ln_modulate_fwd writes it when ld_out >= D + 8), so the
// bias gradient is one more output column of the tensor-core GEMM instead of a separate pass over dy.
// bf16 operands; falls back to GEMM + column-sum kernel when the tcgen05 path does not take the shape.
extern "C" int reed_gemm_wgrad_bias(const void* dy, int64_t ld_dy, const void* x_ext, int64_t ld_x, void* dW, int64_t ldd,
                                    void* db, int n_out, int k_in, int tokens, int accumulate, void* stream) {
  REED_REQUIRE(n_out > 0 && k_in > 0 && tokens > 0 && k_in % 8 == 0 && ld_x >= k_in + 8, "gemm_wgrad_bias: bad shape");
  REED_REQUIRE(db != nullptr, "gemm_wgrad_bias: bias gradient buffer missing");
  EpiParams ep;
  ep.kind = kEpiNone; ep.bias = nullptr; ep.aux = nullptr; ep.ld_aux = 0; ep.gate = nullptr; ep.ld_gate = 0;
  ep.rows_per_group = 1; ep.out2 = nullptr; ep.ld_out2 = 0; ep.accumulate = accumulate;
  ep.bias_grad = (float*)db; ep.n_store = k_in;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = k_in + 8;
  if (tokens > 64 && gemm_tcgen05_supported(ld_dy, ld_x, ldd, dy, x_ext, n_out, N, tokens)) {
    gemm_tcgen05_force_cta_group(0);
    gemm_tcgen05_force_bn(0);
    ++g_tcgen05_launches;
    return gemm_tcgen05(dy, ld_dy, 1, x_ext, ld_x, 1, dW, ldd, kF32, n_out, N, tokens, ep, st);
  }
  ep.bias_grad = nullptr; ep.n_store = 0;
  if (gemm_simt(kBF16, dy, ld_dy, 1, x_ext, ld_x, 1, dW, ldd, kF32, n_out, k_in, tokens, ep, st)) return 1;
  return reed_colsum(dy, kBF16, ld_dy, db, tokens, n_out, stream);
}

// backend: 0 auto (the flash-style tcgen05 kernels when T / head_dim allow, else mma.sync, else SIMT), 1 force SIMT,
// 2 require a tensor-core kernel, 3 require the mma.sync kernel (ragged / short sequences), 4 or 5 require the tcgen05 kernels
static int attn_pick(int act_dtype, int T, int hd, int backend, bool fa_ok, int* which) {
  fa_ok = fa_ok && act_dtype == kBF16;
  const bool mma_ok = act_dtype == kBF16 && attn_mma_supported(T, hd);
  if (backend == 4 || backend == 5) REED_REQUIRE(fa_ok, "attention: tcgen05 path required but T=%d hd=%d unsupported", T, hd);
  if (backend == 3) REED_REQUIRE(mma_ok, "attention: mma.sync path required but T=%d hd=%d unsupported", T, hd);
  if (backend == 2) REED_REQUIRE(fa_ok || mma_ok, "attention: tensor-core path required but T=%d hd=%d unsupported", T, hd);
  if (backend == 1) *which = 0;
  else if (backend == 3) *which = 1;
  else if (fa_ok) *which = 2;
  else if (mma_ok) *which = 1;
  else *which = 0;
  return 0;
}

extern "C" int reed_attn_fwd(int act_dtype, const void* qkv, void* o, void* lse, int B, int T, int H, int hd,
                             int backend, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int which = 0;
  if (attn_pick(act_dtype, T, hd, backend, attn_fa_supported(T, hd), &which)) return 1;
  if (which == 2) return attn_fa_fwd(qkv, o, (float*)lse, B, T, H, hd, st);
  if (which == 1) return attn_mma_fwd(qkv, o, (float*)lse, B, T, H, hd, st);
  return attn_simt_fwd(act_dtype, qkv, o, (float*)lse, B, T, H, hd, st);
}

// delta: [B, H, T] fp32 workspace (dO . O); dqkv: [B, T, 3, H, hd]
extern "C" int reed_attn_bwd(int act_dtype, const void* qkv, const void* o, const void* d_o, const void* lse,
                             void* dqkv, void* delta, int B, int T, int H, int hd, int backend, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int which = 0;
  if (attn_pick(act_dtype, T, hd, backend, attn_fa_bwd_supported(T, hd), &which)) return 1;
  if (which == 2) return attn_fa_bwd(qkv, o, d_o, (const float*)lse, dqkv, (float*)delta, B, T, H, hd, st);
  if (which == 1) return attn_mma_bwd(qkv, o, d_o, (const float*)lse, dqkv, (float*)delta, B, T, H, hd, st);
  return attn_simt_bwd(act_dtype, qkv, o, d_o, (const float*)lse, dqkv, (float*)delta, B, T, H, hd, st);
}
