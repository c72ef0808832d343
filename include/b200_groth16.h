/* libb200groth16.so - C ABI of the B200-native (sm_100a) Groth16 proving backend for davinci-node.
 *
 * This header is the drop-in boundary: a Go shim (davinci-node_b200/go/prover_b200.go, see
 * INTEGRATION.md) binds these symbols with cgo from the reference's `prover` package, replacing
 * the gnark / icicle calls listed beside each entry point.  All file:line citations are into
 * /root/reference (vocdoni/davinci-node).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b200_last_error() returns a
 *     thread-local message.  There is NO CPU fallback anywhere behind this ABI.
 *   - field elements / points use gnark-crypto's in-memory layout (little-endian uint64 limbs,
 *     Montgomery form; G1Affine{X,Y}, G2Affine{X{A0,A1},Y{A0,A1}}, infinity = all-zero), so Go
 *     slices are passed zero-copy as (pointer, length).
 *   - host-pointer functions copy what they need before returning and never retain caller memory
 *     (cgo pointer rule).  `_dev` functions take device pointers and a cudaStream_t (as void*),
 *     enqueue asynchronously and are what a resident pipeline uses.
 */
#ifndef B200_GROTH16_H
#define B200_GROTH16_H

#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* curve ids: the curves dispatched by callGPUProver (prover/prover_gpu.go:24-61) */
enum b200_curve { B200_BN254 = 1, B200_BLS12_377 = 2, B200_BLS12_381 = 3, B200_BW6_761 = 4 };

/* ---- lifecycle ------------------------------------------------------------------------------ */
/* Select GPUs (bit i = CUDA device i; 0 = all visible).  Idempotent and thread-safe.
 * Replaces the icicle runtime/device initialisation done by gnark's icicle backend
 * (imported at prover/prover_gpu.go:10-16). */
B200_API int b200_init(uint32_t device_mask);
B200_API int b200_device_count(void);
B200_API const char* b200_last_error(void);
B200_API const char* b200_version(void);

/* element / point sizes in bytes for a curve (group: 1 = G1, 2 = G2) */
B200_API uint64_t b200_fr_bytes(int curve);
B200_API uint64_t b200_fp_bytes(int curve);
B200_API uint64_t b200_affine_bytes(int curve, int group);
B200_API uint64_t b200_xyzz_bytes(int curve, int group);

/* ---- multi-scalar multiplication ------------------------------------------------------------
 * result = sum_i scalars[i] * points[i], written as ONE affine point (gnark layout).
 * Replaces G1Affine.MultiExp / G2Affine.MultiExp of gnark-crypto, i.e. the `ar`, `bs1`, `Bs`,
 * `krs`, `krs2` computations inside groth16.Prove (prover/prover_cpu.go:37,57) and icicle's
 * Msm / G2Msm (prover/prover_gpu.go:33-56). */
B200_API int b200_msm(int curve, int group, const void* points_affine, const void* scalars_mont, uint64_t n,
             void* out_affine, int device);
/* device-resident variant: out_xyzz receives one extended-Jacobian point {X,Y,ZZ,ZZZ};
 * window_bits = 0 selects the window automatically. */
B200_API int b200_msm_dev(int curve, int group, const void* d_points_affine, const void* d_scalars_mont, uint64_t n,
                 void* d_out_xyzz, int window_bits, void* cuda_stream);
/* normalise `count` XYZZ points to affine on the device */
B200_API int b200_to_affine_dev(int curve, int group, const void* d_xyzz, void* d_affine, uint32_t count, void* cuda_stream);
/* plan the MSM would use for (n, curve): out[0]=window bits c, [1]=windows, [2]=buckets/window,
 * [3]=task size, [4]=group size */
B200_API int b200_msm_plan(int curve, uint64_t n, int window_bits, uint32_t out[5]);

/* ---- debug / parity entry points (device pointers; element-wise over n items) ---------------
 * field: 0 = Fp, 1 = Fr, 2 = Fp2 ; op: 0 add, 1 sub, 2 mul, 3 sqr, 4 from_mont, 5 to_mont, 6 inv, 7 neg */
B200_API int b200_dbg_field_op_dev(int curve, int field, int op, const void* d_a, const void* d_b, void* d_out, uint64_t n,
                          void* cuda_stream);
/* op: 0 madd (xyzz += affine), 1 add (xyzz += xyzz), 2 dbl, 3 to_affine, 4 mul by Fr scalar */
B200_API int b200_dbg_ec_op_dev(int curve, int group, int op, const void* d_a, const void* d_b, void* d_out, uint64_t n,
                       void* cuda_stream);
/* calibration: `iters` dependent Montgomery multiplications per thread over nthreads elements */
B200_API int b200_calib_mul_dev(int curve, int field, void* d_inout, uint64_t nthreads, int iters, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_GROTH16_H */
