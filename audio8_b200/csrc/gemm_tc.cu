// audio8_b200 — host side of the tcgen05 GEMM (tensor maps, tiling, dispatch).  The kernel lives in
// gemm_tc_kernel.cuh and is instantiated in gemm_tc_inst_{kk,kmn,mnmn}.cu.
#include "gemm_tc_kernel.cuh"
#include <stdlib.h>

namespace a8 {
namespace gemm {
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER entry point: it fails with CUDA_ERROR_INVALID_CONTEXT on a host thread that
// has not touched the runtime yet (e.g. autograd's worker thread on its first backward).  Bind the primary context.
void ensure_context() {
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
    cudaFree(nullptr);
    bound = true;
  }
}

int make_tmap(CUtensorMap* out, const a8_operand_t& v, int box0, int box1, const char* what) {
  ensure_context();
  EncodeTiledFn fn = get_encode_fn();
  A8_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  for (int i = 0; i < 4; ++i) {
    A8_REQUIRE(v.dims[i] >= 1, "gemm %s: dim %d is %lld", what, i, (long long)v.dims[i]);
    dims[i] = (cuuint64_t)v.dims[i];
  }
  for (int i = 0; i < 3; ++i) {
    A8_REQUIRE(v.strides[i] > 0 && v.strides[i] % 8 == 0,
               "gemm %s: stride %d = %lld elements is not a positive multiple of 8", what, i + 1,
               (long long)v.strides[i]);
    strides[i] = (cuuint64_t)v.strides[i] * 2ull;
  }
  A8_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15u) == 0, "gemm %s: pointer not 16B aligned", what);
  cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  A8_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled(%s) failed with %d: dims=(%lld,%lld,%lld,%lld) "
             "strides=(%lld,%lld,%lld) box=(%d,%d)",
             what, (int)r, (long long)v.dims[0], (long long)v.dims[1], (long long)v.dims[2],
             (long long)v.dims[3], (long long)v.strides[0], (long long)v.strides[1],
             (long long)v.strides[2], box0, box1);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    // A8_GEMM_SMS=<k>: persistent GEMM grids use k SMs (an even number for CTA pairs).  Data-parallel runs leave a few
    // SMs to the concurrent NCCL all-reduce: a persistent CTA that cannot be scheduled because a collective holds its
    // SM would otherwise start late and stretch the whole launch (tiles are assigned statically).
    if (const char* e = getenv("A8_GEMM_SMS")) {
      const int k = atoi(e);
      if (k >= 2 && k <= n) n = k & ~1;
    }
  }
  return n;
}


int launch_window(const a8_gemm_t& g, KParams& kp, int k16, cudaStream_t stream);  // gemm_tc_window.cu

void copy_coef(OpCoef& o, const a8_operand_t& v) {
  for (int d = 0; d < 4; ++d) {
    o.base[d] = v.base[d]; o.ck[d] = v.ck[d]; o.cb[d] = v.cb[d];
    o.cr[d] = v.cr[d]; o.cl[d] = v.cl[d]; o.ch[d] = v.ch[d];
  }
}

}  // namespace gemm
}  // namespace a8

using namespace a8;
using namespace a8::gemm;

static long long* g_trace = nullptr;
// debug: device buffer of TRACE_CTAS*3*TRACE_TILES*TRACE_EVENTS int64 that the next launches stamp (nullptr = off)
extern "C" void a8_gemm_set_trace(void* buf) { g_trace = static_cast<long long*>(buf); }


extern "C" int a8_gemm(const a8_gemm_t* gp, void* stream_v) {
  A8_REQUIRE(gp != nullptr, "gemm: null descriptor");
  const a8_gemm_t& g = *gp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(g.M > 0 && g.N > 0 && g.k_blocks > 0, "gemm: empty problem M=%d N=%d kb=%d", g.M, g.N,
             g.k_blocks);
  A8_REQUIRE(g.ldc % 8 == 0 && g.c_stride_lo % 8 == 0 && g.c_stride_hi % 8 == 0,
             "gemm: output strides must be multiples of 8 elements");
  A8_REQUIRE(g.c != nullptr && (reinterpret_cast<uintptr_t>(g.c) & 15u) == 0, "gemm: C null or not 16B aligned");
  A8_REQUIRE(g.bias == nullptr || g.N % 8 == 0, "gemm: bias needs N %% 8 == 0");
  A8_REQUIRE(g.bias == nullptr || (reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0, "gemm: bias not 16B aligned");
  A8_REQUIRE(!(g.a.major == MAJOR_MN && g.b.major == MAJOR_K), "gemm: (MN,K) majors not instantiated");
  A8_REQUIRE(g.c_dtype >= OUT_BF16 && g.c_dtype <= OUT_F32_ATOMIC, "gemm: bad c_dtype %d", g.c_dtype);
  int bn = g.block_n;
  if (bn == 0) bn = (g.N <= 64) ? 64 : ((g.N <= 128) ? 128 : 256);
  const int split = g.split_k > 1 ? g.split_k : 1;
  A8_REQUIRE(split <= g.k_blocks, "gemm: split_k %d > k_blocks %d", split, g.k_blocks);
  A8_REQUIRE(split == 1 || g.c_dtype == OUT_F32_ATOMIC, "gemm: split_k needs atomic fp32 output");
  A8_REQUIRE(g.k_inner > 0, "gemm: k_inner must be positive");

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = g.M; kp.N = g.N;
  // a8_gemm_t.reserved doubles as the launch-shape request: low byte 0/1 = plain, 2 = CTA pairs (cta_group::2),
  // 3 = tap-window kernel (gemm_tc_window.cu; bits 8..15 = 16-wide k-steps multiplied per tap, 0 = all four)
  const int mode = g.reserved & 0xFF;
  const int cl = (mode == 2) ? 2 : 1;
  kp.m_tiles = cdiv(cdiv(g.M, BLOCK_M), cl);
  kp.n_tiles = cdiv(g.N, bn);
  kp.lo_count = g.lo_count > 0 ? g.lo_count : 1;
  kp.hi_count = g.hi_count > 0 ? g.hi_count : 1;
  kp.k_blocks = g.k_blocks; kp.k_inner = g.k_inner; kp.split_k = split;
  copy_coef(kp.a, g.a);
  copy_coef(kp.b, g.b);
  kp.c = g.c; kp.c_dtype = g.c_dtype; kp.z_out = g.z_out; kp.aux = g.aux; kp.aux_mode = g.aux_mode;
  kp.bias = g.bias; kp.bias_stride_lo = g.bias_stride_lo; kp.act = g.act; kp.alpha = g.alpha;
  kp.colsum = g.colsum;
  kp.ldc = g.ldc; kp.c_stride_lo = g.c_stride_lo; kp.c_stride_hi = g.c_stride_hi;
  kp.trace = g_trace;
  const long long tiles = (long long)kp.m_tiles * kp.n_tiles * kp.lo_count * kp.hi_count * split;
  A8_REQUIRE(tiles < (1ll << 30), "gemm: too many tiles");
  kp.total_tiles = (int)tiles;

  A8_REQUIRE(g.colsum == nullptr || (g.aux_mode == AUX_MUL && g.c_dtype == OUT_BF16 && split == 1 && kp.lo_count == 1 &&
                                     kp.hi_count == 1 && mode != A8_GEMM_TAP_WINDOW),
             "gemm: colsum needs A8_AUX_MUL, bf16 output, no split-K, one (hi, lo) block and the plain / pair kernel");
  if (mode == A8_GEMM_TAP_WINDOW) return launch_window(g, kp, (g.reserved >> 8) & 0xFF, stream);

  CUtensorMap ma, mb;
  int rc;
  if (g.a.major == MAJOR_K) rc = make_tmap(&ma, g.a, BLOCK_K, BLOCK_M, "A");
  else rc = make_tmap(&ma, g.a, 64, BLOCK_K, "A");
  if (rc) return rc;
  if (g.b.major == MAJOR_K) rc = make_tmap(&mb, g.b, BLOCK_K, bn / cl, "B");
  else rc = make_tmap(&mb, g.b, 64, BLOCK_K, "B");
  if (rc) return rc;

  A8_REQUIRE(g.act >= ACT_NONE && g.act <= ACT_GELU_DZ && g.aux_mode >= AUX_NONE && g.aux_mode <= AUX_MUL, "gemm: bad act / aux_mode");
  const int ek = ek_make(g.c_dtype, g.act, g.z_out != nullptr, g.aux_mode);
  if (g.a.major == MAJOR_K && g.b.major == MAJOR_K) return launch_kk(ek, bn, cl, ma, mb, kp, stream);
  if (g.a.major == MAJOR_K && g.b.major == MAJOR_MN) return launch_kmn(ek, bn, cl, ma, mb, kp, stream);
  return launch_mnmn(ek, bn, cl, ma, mb, kp, stream);
}


// One persistent launch over n problems that share operand majors, coordinate maps, tile shape, split-K factor and
// epilogue kind (see GroupParams in gemm_tc_kernel.cuh).  Used for the weight-gradient GEMMs of the transformer stack:
// every layer's dW = dY^T X (4 per layer) is deferred to the end of the stack's backward and run as one kernel.
namespace {
struct GroupBlob {  // everything one grouped launch needs, built on the host once per set of operand addresses
  GroupParams gp;
  KParams kp;
  int ek, bn, cl, magic;
};
constexpr int GROUP_MAGIC = 0x61386772;
}  // namespace

extern "C" size_t a8_gemm_group_blob_bytes(void) { return sizeof(GroupBlob) + 64; }

extern "C" int a8_gemm_group_launch(const void* blob_v, void* stream_v) {
  A8_REQUIRE(blob_v != nullptr, "gemm_group_launch: null blob");
  const GroupBlob* blob = reinterpret_cast<const GroupBlob*>((reinterpret_cast<uintptr_t>(blob_v) + 63u) & ~uintptr_t(63));
  A8_REQUIRE(blob->magic == GROUP_MAGIC, "gemm_group_launch: blob was not written by a8_gemm_group_prepare");
  return launch_group_mnmn(blob->ek, blob->bn, blob->cl, blob->gp, blob->kp, static_cast<cudaStream_t>(stream_v));
}

extern "C" int a8_gemm_group_prepare(const a8_gemm_t* gs, int32_t n, void* blob_v) {
  A8_REQUIRE(gs != nullptr && n >= 1 && n <= GROUP_MAX, "gemm_group: %d problems (1..%d supported per launch)", n, GROUP_MAX);
  A8_REQUIRE(blob_v != nullptr, "gemm_group_prepare: null blob");
  GroupBlob* blob = reinterpret_cast<GroupBlob*>((reinterpret_cast<uintptr_t>(blob_v) + 63u) & ~uintptr_t(63));
  blob->magic = 0;
  GroupParams& gp = blob->gp;
  KParams& kp = blob->kp;
  const a8_gemm_t& g0 = gs[0];
  int bn = g0.block_n;
  if (bn == 0) bn = 256;
  const int cl = (g0.reserved == 2) ? 2 : 1;
  const int split = g0.split_k > 1 ? g0.split_k : 1;
  A8_REQUIRE(g0.a.major == MAJOR_MN && g0.b.major == MAJOR_MN, "gemm_group: only (MN,MN) operand majors are instantiated");
  A8_REQUIRE(split == 1 || g0.c_dtype == OUT_F32_ATOMIC, "gemm_group: split_k needs atomic fp32 output");
  memset(&kp, 0, sizeof(kp));
  copy_coef(kp.a, g0.a);
  copy_coef(kp.b, g0.b);
  kp.lo_count = kp.hi_count = 1;
  kp.k_inner = g0.k_inner; kp.split_k = split;
  kp.c_dtype = g0.c_dtype; kp.alpha = g0.alpha; kp.aux_mode = AUX_NONE; kp.act = ACT_NONE;
  kp.trace = nullptr;
  long long tiles = 0;
  for (int i = 0; i < n; ++i) {
    const a8_gemm_t& g = gs[i];
    A8_REQUIRE(g.M > 0 && g.N > 0 && g.k_blocks > 0 && g.c != nullptr, "gemm_group[%d]: empty problem", i);
    A8_REQUIRE(g.a.major == g0.a.major && g.b.major == g0.b.major && g.block_n == g0.block_n && g.reserved == g0.reserved &&
                   g.c_dtype == g0.c_dtype && g.split_k == g0.split_k && g.k_inner == g0.k_inner && g.alpha == g0.alpha,
               "gemm_group[%d]: majors / tile shape / output type / split / k_inner / alpha differ from problem 0", i);
    A8_REQUIRE(g.act == ACT_NONE && g.z_out == nullptr && g.aux == nullptr && g.bias == nullptr && g.colsum == nullptr &&
                   (g.lo_count <= 1) && (g.hi_count <= 1),
               "gemm_group[%d]: epilogue extras and batched problems are not supported in groups", i);
    A8_REQUIRE(memcmp(g.a.base, g0.a.base, sizeof(int32_t) * 24) == 0 && memcmp(g.b.base, g0.b.base, sizeof(int32_t) * 24) == 0,
               "gemm_group[%d]: operand coordinate maps differ from problem 0", i);
    A8_REQUIRE(g.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(g.c) & 15u) == 0, "gemm_group[%d]: C alignment", i);
    A8_REQUIRE(split <= g.k_blocks, "gemm_group[%d]: split_k %d > k_blocks %d", i, split, g.k_blocks);
    GroupProb& q = gp.prob[i];
    q.c = g.c; q.ldc = g.ldc; q.M = g.M; q.N = g.N;
    q.m_tiles = cdiv(cdiv(g.M, BLOCK_M), cl);
    q.n_tiles = cdiv(g.N, bn);
    q.k_blocks = g.k_blocks;
    tiles += (long long)q.m_tiles * q.n_tiles * split;
    A8_REQUIRE(tiles < (1ll << 30), "gemm_group: too many tiles");
    q.tile_end = (int)tiles;
    int rc = make_tmap(&gp.map_a[i], g.a, 64, BLOCK_K, "group A");
    if (rc) return rc;
    rc = make_tmap(&gp.map_b[i], g.b, 64, BLOCK_K, "group B");
    if (rc) return rc;
  }
  gp.n_prob = n;
  kp.total_tiles = (int)tiles;
  blob->ek = ek_make(g0.c_dtype, 0, 0, AUX_NONE);
  blob->bn = bn;
  blob->cl = cl;
  blob->magic = GROUP_MAGIC;
  return 0;
}

extern "C" int a8_gemm_group(const a8_gemm_t* gs, int32_t n, void* stream_v) {
  static thread_local GroupBlob* blob = nullptr;
  if (blob == nullptr) blob = static_cast<GroupBlob*>(aligned_alloc(64, (sizeof(GroupBlob) + 63) & ~size_t(63)));
  A8_REQUIRE(blob != nullptr, "gemm_group: out of host memory");
  const int rc = a8_gemm_group_prepare(gs, n, blob);
  return rc ? rc : a8_gemm_group_launch(blob, stream_v);
}
