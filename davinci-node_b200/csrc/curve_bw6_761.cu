// bw6_761 instantiation of the templated sm_100a kernels (see curve_impl.cuh).
#include "curve_impl.cuh"

namespace b200 {

struct Cfg_bw6_761 {
  static constexpr int ID = 4;
  static const char* name() { return "bw6_761"; }
  using Fp = FpT<bw6_761_fp>;
  using Fr = FpT<bw6_761_fr>;
  using G1F = Fp;
  using G2F = FpT<bw6_761_fp>;
};

CurveBackend* backend_bw6_761() {
  static CurveImpl<Cfg_bw6_761> impl;
  return &impl;
}

}  // namespace b200
