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
  using Tower = pairing_bn254;        // extension-field shape of the pairing (pairing.cuh)
  static constexpr int FLAG_BITS = 2;   // gnark-crypto point-compression flag bits (serde.cuh)
  // E: y^2 = x^3 + 3 ; D-twist E': y^2 = x^3 + 3/(9+u)
  static __device__ void curve_b(typename G1F::El& b1, typename G2F::El& b2) {
    typename G1F::El one;
    G1F::set_one(one);
    G1F::mul_small(b1, one, 3);
    typename G2F::El xi;
    G1F::mul_small(xi.c0, one, 9);
    xi.c1 = one;
    G2F::inv(xi, xi);
    G2F::mul_small(b2, xi, 3);
  }
};

CurveBackend* backend_bn254() {
  static CurveImpl<Cfg_bn254> impl;
  return &impl;
}

}  // namespace b200
