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

/* Page-lock / unlock caller memory (cudaHostRegister).  Go heap memory is pageable: copies from it are staged by
 * the driver at a fraction of the PCIe rate (537 MB of witness + constraint vectors per voteverifier proof).  A shim
 * that keeps its solver output buffers across proofs registers them once; b200_prove accepts either kind.
 * Replaces nothing in the reference (gnark's CPU prover never leaves host memory). */
B200_API int b200_host_register(void* ptr, uint64_t bytes);
B200_API int b200_host_unregister(void* ptr);

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
/* Table mode: a base-point set registered once keeps its window multiples T_j[i] = 2^(c j) P_i resident
 * in HBM; MSMs over it need no Horner pass and far fewer buckets (this is how the proving key and the
 * KZG SRS are held; gnark re-reads pk.G1.A etc. on every proof instead).  d_index_map is optional:
 * scalar i multiplies base map[i] (0xffffffff skips it). */
B200_API int b200_bases_create_dev(int curve, int group, const void* d_points_affine, uint64_t n, int window_bits,
                                   uint64_t* handle_out, void* cuda_stream);
B200_API int b200_bases_release(uint64_t handle);
B200_API int b200_msm_bases_dev(uint64_t handle, const void* d_scalars_mont, uint64_t n, const uint32_t* d_index_map,
                                void* d_out_xyzz, void* cuda_stream);
/* plan the MSM would use for (n, curve): out[0]=window bits c, [1]=windows, [2]=buckets/window,
 * [3]=task size, [4]=group size */
B200_API int b200_msm_plan(int curve, uint64_t n, int window_bits, uint32_t out[5]);

/* ---- NTT (gnark-crypto fft.Domain semantics) --------------------------------------------------
 * Replaces fft.Domain.FFT / FFTInverse (gnark-crypto) and icicle's Ntt.
 * A domain holds the twiddle / coset tables resident in HBM for one (curve, size) on the current
 * device.  generator / coset_gen: fr elements in Montgomery form (pk.Domain.Generator,
 * pk.Domain.FrMultiplicativeGen), host pointers. */
B200_API int b200_domain_create(int curve, uint64_t size, const void* generator, const void* coset_gen,
                                uint64_t* domain_out);
B200_API int b200_domain_release(uint64_t domain);
/* in-place transform of `size` fr elements on the device.
 * inverse: 0 FFT, 1 FFTInverse (scaled by 1/n) ; decimation: 0 DIF (natural in, bit-reversed out),
 * 1 DIT (bit-reversed in, natural out) ; coset: 0/1 as gnark's fft.OnCoset() */
B200_API int b200_ntt_dev(uint64_t domain, void* d_data, int inverse, int decimation, int coset, void* cuda_stream);
/* quotient polynomial: d_a <- coefficients of h = (a*b - c)/(X^n - 1) in bit-reversed order, exactly
 * gnark's computeH (7 transforms + fused pointwise work).  d_a, d_b, d_c: `size` fr each (zero padded),
 * all three are overwritten. */
B200_API int b200_compute_h_dev(uint64_t domain, void* d_a, void* d_b, void* d_c, void* cuda_stream);

/* ---- Groth16 proving key / prove ---------------------------------------------------------------
 * b200_slice: (pointer, element count) view of a Go slice; elements in gnark-crypto memory layout. */
typedef struct {
  const void* ptr;
  uint64_t len;
} b200_slice;

/* Everything groth16.Prove reads from *groth16_<curve>.ProvingKey plus the R1CS metadata the
 * prover needs (gnark `ProvingKey` exported fields).  Copied to every selected GPU. */
typedef struct {
  int curve;
  uint64_t domain_size;           /* pk.Domain.Cardinality */
  const void* generator;          /* pk.Domain.Generator            (fr) */
  const void* coset_gen;          /* pk.Domain.FrMultiplicativeGen  (fr) */
  const void* g1_alpha;           /* pk.G1.Alpha, Beta, Delta       (G1Affine) */
  const void* g1_beta;
  const void* g1_delta;
  b200_slice g1_A, g1_B, g1_Z, g1_K;   /* pk.G1.A / B / Z / K */
  const void* g2_beta;            /* pk.G2.Beta, Delta              (G2Affine) */
  const void* g2_delta;
  b200_slice g2_B;                /* pk.G2.B */
  b200_slice infinity_a;          /* pk.InfinityA, pk.InfinityB: []bool, one byte per wire */
  b200_slice infinity_b;
  uint64_t nb_wires;              /* len(InfinityA) = internal + secret + public variables */
  uint64_t nb_public;             /* r1cs.GetNbPublicVariables() (includes the constant-one wire) */
  b200_slice krs_skip;            /* ascending uint32 wire ids left out of the K MSM: every
                                     commitmentInfo[i].PrivateCommitted wire and CommitmentIndex wire */
  uint32_t nb_commitments;        /* len(pk.CommitmentKeys) */
  const b200_slice* commit_basis;            /* pk.CommitmentKeys[i].Basis */
  const b200_slice* commit_basis_exp_sigma;  /* pk.CommitmentKeys[i].BasisExpSigma */
  uint64_t z_offset;              /* 0 for a whole key.  Range-split keys (one slice of the key per GPU,
                                     SURVEY.md 8e-2): g1_Z holds Z[z_offset .. z_offset+len) and pairs with
                                     h[z_offset ..]; the wire-indexed arrays are sliced by the caller */
} b200_pk_desc;

typedef struct {
  b200_slice wires;               /* solution.W : nb_wires fr */
  b200_slice a, b, c;             /* solution.A / B / C : nbConstraints fr each */
  const void* r;                  /* prover randomness, fr Montgomery (pinned by tests, crypto/rand otherwise) */
  const void* s;
  uint32_t nb_commitments;
  const b200_slice* priv_committed;   /* privateCommittedValues[i] */
  const void* fold_challenge;     /* fr; only read when nb_commitments > 1 */
  uint32_t abc_form;              /* 0: a, b, c are the solver's constraint values (the usual case).
                                     1: a, b, c are already their evaluations on the coset g<omega>, domain_size
                                        elements each (b200_pk_coset_evals_dev): range-split proving shards the
                                        quotient - each GPU transforms one vector and broadcasts it over NVLink */
} b200_prove_in;

typedef struct {
  void* ar;    /* proof.Ar   G1Affine */
  void* bs;    /* proof.Bs   G2Affine */
  void* krs;   /* proof.Krs  G1Affine */
  void* pok;   /* proof.CommitmentPok G1Affine (untouched when the circuit has no commitment) */
} b200_proof_out;

/* Upload a proving key (replaces icicle's device-side pk setup behind prover/prover_gpu.go:24-61). */
B200_API int b200_pk_register(const b200_pk_desc* desc, uint64_t* handle_out);
B200_API int b200_pk_release(uint64_t handle);
/* HBM budget policy.  A key is held as window tables T_q[i] = 2^(c s q) P_i; the table stride s trades HBM for
 * bucket-reduction work (s = 1: ~23 GB for a 2^22 BLS12-377 key, s = 2: half of that).  Registration picks the
 * smallest s whose tables fit `bytes` (0 = automatic: 55% of the device memory free at registration, so the three
 * production keys loaded one after the other by sequencer/circuit_artifacts.go:39-76 co-reside on one B200).
 * b200_pk_info: out[0] = table stride, [1] = resident table bytes on the first GPU, [2] = proofs in flight per GPU,
 * [3] = window bits of the wire-indexed sets, [4] = window bits of Z, [5] = number of GPUs holding the key. */
B200_API int b200_set_pk_table_budget(uint64_t bytes);
B200_API int b200_pk_info(uint64_t handle, uint64_t out[6]);
/* Pedersen commitment i over Basis: called from the BSB22 solver hint. */
B200_API int b200_commit(uint64_t handle, uint32_t i, b200_slice values, void* out_g1_affine, int device);
/* The proof: replaces groth16.Prove's computeH + MultiExp section (prover/prover_cpu.go:37,57;
 * gpugroth16.Prove at prover/prover_gpu.go:33-56).  device = -1 picks a GPU round-robin. */
B200_API int b200_prove(uint64_t handle, const b200_prove_in* in, const b200_proof_out* out, int device);
/* same, every pointer in `in` / `out` is a device pointer on the key's device (resident pipeline) */
B200_API int b200_prove_dev(uint64_t handle, const b200_prove_in* in, const b200_proof_out* out, int device);

/* Range-split proving (one large proof over several GPUs).  Every GPU holds a slice of the
 * key (b200_pk_register with a sliced descriptor) and runs b200_prove_partial_dev on its slice of the wire
 * vector and the FULL a, b, c: it writes its six un-normalised partial sums
 *   [Ar, Bs1, K, Z, Pok (G1 XYZZ each), Bs (G2 XYZZ)]  =  5 * xyzz_bytes(1) + xyzz_bytes(2) bytes.
 * Every slice folds s*Ar_g + r*Bs1_g into its K partial (sum_g s*Ar_g = s*Ar), so the two scalar multiplications
 * of the assembly run on every GPU beside its MSMs instead of after the gather.
 * The caller all-gathers the partials over NVLink (NCCL cannot reduce with the group law) and any GPU finishes
 * with b200_assemble_dev, which adds the `nparts` partials per element and normalises to affine.  d_r / d_s are
 * ignored (kept for ABI stability).  d_out: Ar | Krs | Pok (G1 affine) | Bs (G2 affine).
 * Device inputs of every `_dev` prove entry point are read on the library's own streams: they must be complete
 * (producer stream synchronised) before the call. */
B200_API int b200_prove_partial_dev(uint64_t handle, const b200_prove_in* in, void* d_partials_out, int device);
/* First half of gnark's computeH for ONE vector: d_vec (domain_size fr on the key's device, natural
 * order, zero padded) <- its evaluations on the coset g<omega> (FFTInverse DIF, then FFT DIT OnCoset), asynchronously on
 * cuda_stream.  With the three vectors transformed (on three different GPUs and broadcast), b200_prove*_dev with
 * abc_form = 1 runs only the pointwise division and the last inverse transform. */
B200_API int b200_pk_coset_evals_dev(uint64_t handle, void* d_vec, int device, void* cuda_stream);
B200_API int b200_assemble_dev(int curve, const void* d_partials, uint32_t nparts, const void* d_r, const void* d_s,
                               int have_pok, void* d_out, void* cuda_stream);

/* ---- EIP-4844 blob commitment (BLS12-381) -----------------------------------------------------
 * Replaces gethkzg.BlobToCommitment (types/blobs.go:90-96). */
/* g1_lagrange: npoints x 48-byte compressed points in the order of config/kzg_trusted_setup.txt */
B200_API int b200_kzg_srs_register(const uint8_t* g1_lagrange, uint32_t npoints, uint64_t* handle_out);
B200_API int b200_kzg_srs_release(uint64_t handle);
B200_API int b200_blob_commit(uint64_t srs, const uint8_t* blob /* npoints*32 bytes */, uint8_t commitment_out[48],
                              int device);
/* KZG opening proof of the blob polynomial at `point` (32 big-endian bytes, canonical): proof_out = 48-byte
 * compressed commitment to (p(X) - p(z)) / (X - z), claim_out = y = p(z) as 32 big-endian bytes.  Replaces
 * gethkzg.ComputeProof (types/blobs.go:123-134); with point = the Fiat-Shamir challenge
 * sha256("FSBLOBVERIFY_V1_" || u128(4096) || blob || commitment) mod r, computed by the host shim, it is
 * gethkzg.ComputeBlobProof (types/blobs.go:111-117).  Points inside the evaluation domain are handled as in
 * EIP-4844 compute_quotient_eval_within_domain. */
/* EIP-7594 cell proofs.  b200_kzg_srs_add_monomial registers the ceremony's monomial-basis points [tau^j]_1
 * (npoints x 48-byte compressed, the third block of config/kzg_trusted_setup.txt); b200_blob_cell_proofs then writes
 * the 128 x 48-byte proofs of the blob's cells: inverse NTT to coefficients, the 128 quotients by X^64 - h_k^64 and
 * 128 MSMs on the GPU.  Replaces gethkzg.ComputeCellProofs (types/blobs.go:99-105). */
B200_API int b200_kzg_srs_add_monomial(uint64_t srs, const uint8_t* g1_monomial, uint32_t npoints);
B200_API int b200_blob_cell_proofs(uint64_t srs, const uint8_t* blob, uint8_t* proofs_out /* 128*48 bytes */, int device);
B200_API int b200_blob_proof(uint64_t srs, const uint8_t* blob, const uint8_t point_be[32], uint8_t proof_out[48],
                             uint8_t claim_out[32], int device);

/* ---- proving-key artefacts ----------------------------------------------------------------------------
 * gnark-crypto compressed points -> affine points in memory layout, on the device, for any curve and group: the bulk of
 * loading a proving key written by pk.WriteTo (cmd/circuit-compile/main.go:507-512) and read at
 * circuits/artifacts.go:391-406 (pk.UnsafeReadFrom: ~10^7 square roots on the CPU in the reference).  Encoding: big-endian
 * X (G2 over Fp2: X.A1 || X.A0), flags in the top bits of byte 0 (BN254: 2 bits, the others 3).
 * d_err_flags (one uint32, zeroed by the caller) gets bit 0 = bad encoding, bit 1 = not on the curve.  compress is the
 * inverse (pk.WriteTo). */
B200_API uint64_t b200_compressed_bytes(int curve, int group);
B200_API int b200_points_decompress_dev(int curve, int group, const void* d_bytes, uint64_t n, void* d_affine_out,
                                        uint32_t* d_err_flags, void* cuda_stream);
B200_API int b200_points_compress_dev(int curve, int group, const void* d_affine, uint64_t n, void* d_bytes_out,
                                      void* cuda_stream);

/* ---- proof verification: product-of-pairings check ----------------------------------------------------
 * *result_out = 1 when prod_i e(P_i, Q_i) == 1, else 0.  This is the pairing check of gnark's groth16.Verify, which the
 * reference runs right after every proof (circuits/artifacts.go:595-613: e(-Ar, Bs) e(alpha, beta) e(L, gamma) e(Krs, delta)
 * == 1, and the Pedersen check e(C, GSigmaNeg) e(PoK, G) == 1), i.e. gnark-crypto's <curve>.PairingCheck, and of the EIP-197
 * precompile call in config/statetransition_vkey.sol:720-746.  Points: n affine G1 / G2 points in gnark memory layout
 * (Montgomery limbs; all-zero = infinity), HOST pointers.  The reduced Tate pairing is used (csrc/pairing.cuh): the
 * predicate equals gnark's, the group element does not (optimal ate differs by a fixed exponent).  gt_out (optional,
 * b200_gt_bytes(curve) bytes): the value of the product as k Fp coefficients of Fp[w]/(w^k + ...), Montgomery limbs.
 * Fails (non-zero status) when some P_i is not in the order-r subgroup.  The caller owns the rest of Verify: the
 * commitment challenges (hash-to-field) and the public-input MSM (b200_msm). */
B200_API uint64_t b200_gt_bytes(int curve);
B200_API int b200_pairing_check(int curve, const void* g1_affine, const void* g2_affine, uint32_t n, int* result_out,
                                void* gt_out, int device);
/* Throughput form: n_checks independent checks of pairs_per_check pairs each (check c uses pairs [c * per, (c + 1) * per)),
 * one GPU thread per pair and per check - batches of proofs (sequencer/aggregate.go:503-519 re-verifies every inner proof
 * of a batch; api/workers.go:353 every worker submission).  results_out[c] = 1 holds, 0 does not, -1 some G1 point of
 * the check is outside the order-r subgroup. */
B200_API int b200_pairing_check_batch(int curve, const void* g1_affine, const void* g2_affine, uint32_t pairs_per_check,
                                      uint32_t n_checks, int32_t* results_out, int device);

/* ---- setup building block / instrumentation -----------------------------------------------------
 * out[i] = [k_i] base as affine points: the fixed-base batch scalar multiplication groth16.Setup is
 * made of (prover/setup.go:15-28 -> groth16.Setup).  Device pointers. */
B200_API int b200_fixed_base_dev(int curve, int group, const void* d_base_affine, const void* d_scalars_mont, uint64_t n,
                                 void* d_out_affine, void* cuda_stream);
/* Combine step of a range-split MSM: out = affine(sum of `count` XYZZ partial sums, one
 * per GPU, all-gathered over NVLink by the caller).  Replaces nothing in the reference (it has no
 * multi-GPU path); gnark sums its per-chunk partials the same way on the CPU. */
B200_API int b200_sum_partials_dev(int curve, int group, const void* d_xyzz, uint32_t count, void* d_out_affine,
                                   void* cuda_stream);
/* number of this library's kernels launched so far in the process */
B200_API uint64_t b200_launch_count(void);
/* CUDA-event kernel timers: tags 0 = G1 bucket accumulation, 1 = G2 bucket accumulation, 2 = NTT pass,
 * 3 = whole G1 MSM, 4 = whole G2 MSM.  collect() synchronises, sums ms / launch counts per tag, clears. */
B200_API int b200_profile_enable(int on);
B200_API int b200_profile_collect(double ms_out[5], uint64_t count_out[5]);
/* every instrumented phase recorded since b200_profile_enable(1), 4 doubles per record: [tag, stream ordinal,
 * start ms, end ms].  Further tags: 5 digit/sort, 6 bucket schedule, 7 oversized buckets, 8 bucket reduction,
 * 9 window sums, 10 input copies, 11 proof assembly, 12 affine pre-reduction (opt-in).  Synchronises the device;
 * does not clear. */
B200_API int b200_profile_timeline(double* out, uint64_t cap_records, uint64_t* n_out);

/* ---- debug / parity entry points (device pointers; element-wise over n items) ---------------
 * field: 0 = Fp, 1 = Fr, 2 = Fp2 ; op: 0 add, 1 sub, 2 mul, 3 sqr, 4 from_mont, 5 to_mont, 6 inv, 7 neg, 8 sqrt (zero when not a square), 9 inversion by binary GCD */
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
