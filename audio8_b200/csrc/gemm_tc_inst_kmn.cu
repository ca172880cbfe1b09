// audio8_b200 — tcgen05 GEMM instantiations for operand majors (MAJOR_K, MAJOR_MN); see gemm_tc_kernel.cuh.
#include "gemm_tc_kernel.cuh"

namespace a8 {
namespace gemm {

int launch_kmn(int ek, int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t s) {
  switch (ek) {
    case ek_make(OUT_BF16, 0, 0, AUX_NONE): return launch_bn<MAJOR_K, MAJOR_MN, ek_make(OUT_BF16, 0, 0, AUX_NONE)>(bn, cl, ma, mb, kp, s);
    case ek_make(OUT_BF16, 0, 0, AUX_ADD): return launch_bn<MAJOR_K, MAJOR_MN, ek_make(OUT_BF16, 0, 0, AUX_ADD)>(bn, cl, ma, mb, kp, s);
    case ek_make(OUT_BF16, 0, 0, AUX_MUL): return launch_bn<MAJOR_K, MAJOR_MN, ek_make(OUT_BF16, 0, 0, AUX_MUL)>(bn, cl, ma, mb, kp, s);
    case ek_make(OUT_F32, 0, 0, AUX_NONE): return launch_bn<MAJOR_K, MAJOR_MN, ek_make(OUT_F32, 0, 0, AUX_NONE)>(bn, cl, ma, mb, kp, s);
  }
  return launch_bn<MAJOR_K, MAJOR_MN, EK_GENERIC>(bn, cl, ma, mb, kp, s);
}

}  // namespace gemm
}  // namespace a8
