// bn254 instantiation of the templated sm_100a kernels (see curve_impl.cuh).
#include "curve_impl.cuh"

namespace b200 {

struct Cfg_bn254 {
  static constexpr int ID = 1;
  static const char* name() { return "bn254"; }
  using Fp = FpT<bn254_fp>;
  using Fr = FpT<bn254_fr>;
  using G1F = Fp;
  using G2F = Fp2T<bn254_fp, 1>;
};

CurveBackend* backend_bn254() {
  static CurveImpl<Cfg_bn254> impl;
  return &impl;
}

}  // namespace b200
