// bls12_377 instantiation of the templated sm_100a kernels (see curve_impl.cuh).
#include "curve_impl.cuh"

namespace b200 {

struct Cfg_bls12_377 {
  static constexpr int ID = 2;
  static const char* name() { return "bls12_377"; }
  using Fp = FpT<bls12_377_fp>;
  using Fr = FpT<bls12_377_fr>;
  using G1F = Fp;
  using G2F = Fp2T<bls12_377_fp, 5>;
};

CurveBackend* backend_bls12_377() {
  static CurveImpl<Cfg_bls12_377> impl;
  return &impl;
}

}  // namespace b200
