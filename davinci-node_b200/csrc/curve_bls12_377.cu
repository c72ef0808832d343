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
  using Tower = pairing_bls12_377;        // extension-field shape of the pairing (pairing.cuh)
  static constexpr int FLAG_BITS = 3;   // gnark-crypto point-compression flag bits (serde.cuh)
  // E: y^2 = x^3 + 1 ; D-twist E': y^2 = x^3 + 1/u
  static __device__ void curve_b(typename G1F::El& b1, typename G2F::El& b2) {
    G1F::set_one(b1);
    typename G2F::El u;
    G1F::set_zero(u.c0);
    G1F::set_one(u.c1);
    G2F::inv(b2, u);
  }
};

CurveBackend* backend_bls12_377() {
  static CurveImpl<Cfg_bls12_377> impl;
  return &impl;
}

}  // namespace b200
