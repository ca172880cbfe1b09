// audio8_b200 — tcgen05 GEMM instantiations for operand majors (MAJOR_MN, MAJOR_MN); see gemm_tc_kernel.cuh.
#include "gemm_tc_kernel.cuh"

namespace a8 {
namespace gemm {

int launch_mnmn(int ek, int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t s) {
  switch (ek) {
    case ek_make(OUT_F32, 0, 0, AUX_NONE): return launch_bn<MAJOR_MN, MAJOR_MN, ek_make(OUT_F32, 0, 0, AUX_NONE)>(bn, cl, ma, mb, kp, s);
    case ek_make(OUT_F32_ATOMIC, 0, 0, AUX_NONE): return launch_bn<MAJOR_MN, MAJOR_MN, ek_make(OUT_F32_ATOMIC, 0, 0, AUX_NONE)>(bn, cl, ma, mb, kp, s);
  }
  return launch_bn<MAJOR_MN, MAJOR_MN, EK_GENERIC>(bn, cl, ma, mb, kp, s);
}

// grouped launches (a8_gemm_group): weight-gradient groups only, 256 x 256 pair tiles (or 128-wide for narrow outputs)
int launch_group_mnmn(int ek, int bn, int cl, const GroupParams& gp, const KParams& kp, cudaStream_t s) {
  if (cl == 2 && bn == 256) {
    if (ek == ek_make(OUT_F32, 0, 0, AUX_NONE))
      return launch_group_inst<MAJOR_MN, MAJOR_MN, 256, 2, ek_make(OUT_F32, 0, 0, AUX_NONE)>(gp, kp, s);
    if (ek == ek_make(OUT_F32_ATOMIC, 0, 0, AUX_NONE))
      return launch_group_inst<MAJOR_MN, MAJOR_MN, 256, 2, ek_make(OUT_F32_ATOMIC, 0, 0, AUX_NONE)>(gp, kp, s);
  }
  if (cl == 1 && bn == 128 && ek == ek_make(OUT_F32, 0, 0, AUX_NONE))
    return launch_group_inst<MAJOR_MN, MAJOR_MN, 128, 1, ek_make(OUT_F32, 0, 0, AUX_NONE)>(gp, kp, s);
  set_error("gemm_group: (MN,MN) groups are instantiated for block_n 256 pairs (fp32 store / reduce) and 128 single (fp32 store); "
            "got block_n %d cluster %d epilogue kind %d", bn, cl, ek);
  return -1;
}

}  // namespace gemm
}  // namespace a8
