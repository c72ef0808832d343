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
  using Tower = pairing_bls12_381;        // extension-field shape of the pairing (pairing.cuh)
  static constexpr int FLAG_BITS = 3;   // gnark-crypto point-compression flag bits (serde.cuh)
  // E: y^2 = x^3 + 4 ; M-twist E': y^2 = x^3 + 4(1+u)
  static __device__ void curve_b(typename G1F::El& b1, typename G2F::El& b2) {
    typename G1F::El one;
    G1F::set_one(one);
    G1F::mul_small(b1, one, 4);
    b2.c0 = b1;
    b2.c1 = b1;
  }
};

CurveBackend* backend_bls12_381() {
  static CurveImpl<Cfg_bls12_381> impl;
  return &impl;
}

}  // namespace b200
