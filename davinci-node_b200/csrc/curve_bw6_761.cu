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
  using Tower = pairing_bw6_761;        // extension-field shape of the pairing (pairing.cuh)
  static constexpr int FLAG_BITS = 3;   // gnark-crypto point-compression flag bits (serde.cuh)
  // E: y^2 = x^3 - 1 ; M-twist E' (over Fp): y^2 = x^3 + 4
  static __device__ void curve_b(typename G1F::El& b1, typename G2F::El& b2) {
    typename G1F::El one;
    G1F::set_one(one);
    G1F::neg(b1, one);
    G2F::mul_small(b2, one, 4);
  }
};

CurveBackend* backend_bw6_761() {
  static CurveImpl<Cfg_bw6_761> impl;
  return &impl;
}

}  // namespace b200
