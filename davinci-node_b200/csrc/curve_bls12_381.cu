// bls12_381 instantiation of the templated sm_100a kernels (see curve_impl.cuh).
#include "curve_impl.cuh"

namespace b200 {

struct Cfg_bls12_381 {
  static constexpr int ID = 3;
  static const char* name() { return "bls12_381"; }
  using Fp = FpT<bls12_381_fp>;
  using Fr = FpT<bls12_381_fr>;
  using G1F = Fp;
  using G2F = Fp2T<bls12_381_fp, 1>;
};

CurveBackend* backend_bls12_381() {
  static CurveImpl<Cfg_bls12_381> impl;
  return &impl;
}

}  // namespace b200
