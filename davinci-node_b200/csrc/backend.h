// Host-side internal interface between the C-ABI (capi.cu) and the per-curve template
// instantiations (curve_*.cu).  Not part of the public boundary (that is include/b200_groth16.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define B200_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                              std::to_string(__LINE__) + ")");                                       \
  } while (0)

// ---- instrumentation (off by default): kernel launch counter and CUDA-event kernel timers
enum ProfTag { PROF_MSM_ACC_G1 = 0, PROF_MSM_ACC_G2 = 1, PROF_NTT_PASS = 2, PROF_MSM_TOTAL_G1 = 3, PROF_MSM_TOTAL_G2 = 4,
               PROF_TAGS = 5,   // tags below only appear in the timeline (b200_profile_timeline)
               PROF_MSM_SORT = 5, PROF_MSM_SCHED = 6, PROF_MSM_OVF = 7, PROF_MSM_BUCKET_REDUCE = 8, PROF_MSM_SUMS = 9,
               PROF_INPUTS = 10, PROF_ASSEMBLE = 11, PROF_MSM_PRE = 12, PROF_ALL_TAGS = 13 };
void prof_count_launches(uint64_t n);
uint64_t prof_launches();
bool prof_enabled();
// returns an opaque token (or -1 when profiling is off); records the start event on `s`
int prof_begin(int tag, cudaStream_t s);
void prof_end(int token, cudaStream_t s);

// Growable device scratch buffer (stream-ordered use; never shrinks).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) B200_CUDA(cudaFree(p));
      p = nullptr;
      cap = 0;
      size_t want = bytes + (bytes >> 3) + 256;
      B200_CUDA(cudaMalloc(&p, want));
      cap = want;
    }
    return p;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Scratch for one MSM in flight (one per stream).  The latency-bound tail of an MSM (oversized-bucket
// merge, bucket reduction, window sums: few blocks, long dependent chains) is enqueued on `tail`, a
// highest-priority stream, so its blocks are dispatched ahead of the thousands of queued accumulation
// blocks of whatever bulk kernel shares the GPU; ordering with the caller's stream is kept by events.
struct MsmWorkspace {
  DevBuf hist, off, cur, sorted, chunk_sums, buckets, tasks, obuckets, partial, mid, groups, windows, ctr, perm, bins;
  DevBuf pre_a, pre_b, pre_prefix, pre_tot, off2;   // affine pre-reduction (msm_pre.cuh)
  cudaStream_t tail = nullptr;
  cudaEvent_t e_fwd = nullptr, e_back = nullptr;
  void ensure_tail() {
    if (tail) return;
    int lo = 0, hi = 0;
    B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi is the numerically lowest = greatest priority
    B200_CUDA(cudaStreamCreateWithPriority(&tail, cudaStreamNonBlocking, hi));
    B200_CUDA(cudaEventCreateWithFlags(&e_fwd, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&e_back, cudaEventDisableTiming));
  }
  // work enqueued on `tail` after hop_to_tail(s) runs after everything already in s; hop_back makes s wait for it
  void hop_to_tail(cudaStream_t s) {
    ensure_tail();
    B200_CUDA(cudaEventRecord(e_fwd, s));
    B200_CUDA(cudaStreamWaitEvent(tail, e_fwd, 0));
  }
  void hop_back(cudaStream_t s, bool join) {   // !join: only mark the end of the tail (see wait_tail)
    B200_CUDA(cudaEventRecord(e_back, tail));
    if (join) B200_CUDA(cudaStreamWaitEvent(s, e_back, 0));   // (s may be the legacy default stream, 0)
  }
  void wait_tail(cudaStream_t s) { B200_CUDA(cudaStreamWaitEvent(s, e_back, 0)); }
  MsmWorkspace() = default;
  MsmWorkspace(const MsmWorkspace&) = delete;
  MsmWorkspace& operator=(const MsmWorkspace&) = delete;
  ~MsmWorkspace() {
    if (tail) cudaStreamDestroy(tail);
    if (e_fwd) cudaEventDestroy(e_fwd);
    if (e_back) cudaEventDestroy(e_back);
  }
};

// A base-point set with its precomputed window multiples resident in HBM ("table mode", msm_plan.h)
struct MsmBases {
  DevBuf tables;      // ntab tables of npts affine points: T_q[i] = 2^(c tstride q) P_i
  uint64_t npts = 0;
  int c = 0, nwin = 0, group = 1;
  int tstride = 1;    // table stride (msm_plan.h): 1 = a table per digit window
  int ntab = 0;       // ceil(nwin / tstride)
};

}  // namespace b200
#include "msm_plan.h"
namespace b200 {

// Output of the digit / counting-sort stage (device pointers into the sorting workspace): entries of
// bucket array a live at sorted[a * pl.stride + off[a * nb + b] .. end[a * nb + b])
struct MsmSorted {
  MsmPlan pl;
  uint32_t* off = nullptr;
  uint32_t* end = nullptr;
  uint32_t* sorted = nullptr;
  uint32_t* totals = nullptr;
};

struct MsmStats {
  int c, nwin;
  uint32_t nb, task, group;
};

enum FieldSel { FIELD_FP = 0, FIELD_FR = 1, FIELD_FP2 = 2 };
enum FieldOp { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SQR = 3, OP_FROM_MONT = 4, OP_TO_MONT = 5, OP_INV = 6, OP_NEG = 7,
               OP_SQRT = 8,     // sqrt: a root, or zero when the input is not a square
               OP_INV_BIN = 9 };  // binary-GCD inversion (the one-thread path of the proof assembly)
enum EcOp { EC_MADD = 0, EC_ADD = 1, EC_DBL = 2, EC_TO_AFFINE = 3, EC_MUL_SCALAR = 4 };

// Device-resident evaluation domain of size 2^logn (twiddle / coset-power tables live in HBM).
struct NttDomain {
  int logn = 0;
  int lo_bits = 0;      // split of the two-level coset power tables
  DevBuf tw_fwd, tw_inv;               // omega^k, omega^-k   (k < n/2)
  DevBuf g_lo, g_hi, g_hi_scaled;      // g^j ; g^(j 2^lo_bits) ; same times 1/n
  DevBuf gi_lo, gi_hi_scaled;          // g^-j ; g^-(j 2^lo_bits) / n
  DevBuf consts;                       // [omega^-1, g^-1, 1/n, 1/(g^n-1), omega, g]
};

// Inputs of the final proof assembly (all device pointers, XYZZ points unless noted).
struct AssembleArgs {
  const void* ar_msm;     // sum w_i A_i + alpha + r delta      (G1)
  const void* bs1_msm;    // sum w_i B_i + beta + s delta       (G1)
  const void* bs2_msm;    // same in G2
  const void* k_msm;      // sum w_i K_i - r s delta            (G1; written by phase 4)
  const void* z_msm;      // sum h_i Z_i                        (G1)
  const void* pok_msm;    // may be null                        (G1)
  const void* rs;         // 4 Fr elements (Montgomery): r, s, 1, -r s
  void* tmp;              // 2 G1 XYZZ scratch
  void* out_ar;           // G1 affine
  void* out_bs;           // G2 affine
  void* out_krs;          // G1 affine
  void* out_pok;          // G1 affine (if pok_msm)
};

// One implementation per curve (template instantiation in curve_<name>.cu).
struct CurveBackend {
  virtual ~CurveBackend() = default;
  virtual int id() const = 0;
  virtual const char* name() const = 0;
  virtual size_t fr_bytes() const = 0;
  virtual size_t fp_bytes() const = 0;
  virtual size_t affine_bytes(int group) const = 0;   // group: 1 = G1, 2 = G2
  virtual size_t xyzz_bytes(int group) const = 0;
  virtual int fr_bits() const = 0;

  // --- MSM: device pointers, result (one XYZZ point) written to d_out; fully asynchronous on `s`
  // d_index_map (optional): scalar i multiplies points[map[i]]; map[i] == 0xffffffff skips scalar i
  // join = false: `s` does not wait for the MSM's latency-bound tail (which runs on ws.tail); the result is
  // ready for any stream that calls ws.wait_tail(stream)
  virtual void msm(int group, const void* d_points, const void* d_scalars, uint64_t n, void* d_out_xyzz,
                   MsmWorkspace& ws, cudaStream_t s, int c_override, MsmStats* stats,
                   const uint32_t* d_index_map = nullptr, const MsmBases* bases = nullptr, bool join = true,
                   int pre = -1) = 0;   // pre: affine pre-reduction levels (table mode), -1 = default
  // Shared-scalar MSMs (table mode): ONE digit/sort pass over `n` scalars feeds up to 4 base sets, set j
  // taking point maps[j][i] for scalar i (0xffffffff = skip).  msm_reduce then accumulates and reduces
  // `count` consecutive sets (their tables in `bases`, all of `group`) into `count` XYZZ results; a G1 and a
  // G2 key over the same wires (Groth16's B) reduce the same sorted set twice.
  virtual void msm_sort(const void* d_scalars, uint64_t n, const MsmBases* const* bases, const uint32_t* const* maps,
                        int nsets, MsmWorkspace& ws, cudaStream_t s, MsmSorted& out) = 0;
  virtual void msm_reduce(int group, const MsmSorted& so, int first_set, int count, const MsmBases* const* bases,
                          void* d_out_xyzz, MsmWorkspace& ws, cudaStream_t s, bool join = true) = 0;
  // Batched MSM: `batch` scalar vectors of n_per scalars each over ONE base set -> `batch` XYZZ results, with one
  // sort and one accumulate / reduce pass for all of them (the 128 quotients of the EIP-7594 cell proofs)
  virtual void msm_batch(int group, const MsmBases& bases, const void* d_scalars, uint64_t n_per, uint32_t batch,
                         const uint32_t* d_index_map, void* d_out_xyzz, MsmWorkspace& ws, cudaStream_t s) = 0;
  // window width the cost model picks for a table-mode base set of npts points
  virtual int table_window(uint64_t npts, int tstride = 1) const = 0;
  // precompute T_q[i] = 2^(c tstride q) P_i for a base set (window_bits = 0: cost model)
  virtual void build_tables(MsmBases& b, int group, const void* d_points, uint64_t npts, int window_bits,
                            cudaStream_t s, int tstride = 1) = 0;
  // --- NTT / quotient
  virtual void domain_init(NttDomain& d, int logn, const void* d_omega, const void* d_g, cudaStream_t s) = 0;
  // gnark fft.Domain semantics: inverse ? FFTInverse : FFT ; dit ? DIT (bit-reversed in) : DIF ; coset
  virtual void ntt(NttDomain& d, void* d_data, bool inverse, bool dit, bool coset, cudaStream_t s) = 0;
  // a <- h = ((a*b - c) / Z) coefficients in bit-reversed order (gnark computeH); a, b, c natural order, length n
  virtual void compute_h(NttDomain& d, void* d_a, void* d_b, void* d_c, cudaStream_t s) = 0;
  // the two halves of compute_h, for a quotient sharded over GPUs (range-split proving): coset_evals turns ONE of
  // a / b / c (natural order, zero padded) into its evaluations on the coset g<omega> (inverse DIF, coset DIT);
  // compute_h_tail takes the three evaluation vectors to h (pointwise (a*b - c)/(g^n - 1), inverse coset DIF)
  virtual void coset_evals(NttDomain& d, void* d_v, cudaStream_t s) = 0;
  virtual void compute_h_tail(NttDomain& d, void* d_a, void* d_b, void* d_c, cudaStream_t s) = 0;
  virtual void scale_vec(void* d_x, const void* d_c, uint64_t n, cudaStream_t s) = 0;
  // writes r, s, 1, -r*s (Montgomery) to d_out[0..4) from canonical-or-Montgomery inputs already on device
  virtual void prep_rs(const void* d_r, const void* d_s, void* d_out4, cudaStream_t s) = 0;
  // phases (bit mask): 1 = the two scalar multiplications tmp = {s*Ar, r*Bs1} (need only the wire-indexed G1 MSMs),
  // 2 = sums and affine normalisation (needs everything), 4 = k_msm += tmp[0] + tmp[1] in place (range-split slices)
  virtual void assemble(const AssembleArgs& a, cudaStream_t s, int phases = 3) = 0;
  // --- EIP-4844 helpers (BLS12-381 only; other curves throw)
  virtual void g1_decompress(const void* d_bytes, void* d_affine, uint32_t n, uint32_t* d_err, cudaStream_t s) = 0;
  virtual void g1_compress(const void* d_affine, void* d_bytes, uint32_t n, cudaStream_t s) = 0;
  // gnark-crypto compressed points of any curve / group (serde.cuh): the on-disk format of proving keys.
  // d_err: bit 0 bad encoding, bit 1 not on the curve
  virtual size_t compressed_bytes(int group) const = 0;
  virtual void points_decompress(int group, const void* d_bytes, void* d_affine, uint64_t n, uint32_t* d_err, cudaStream_t s) = 0;
  virtual void points_compress(int group, const void* d_affine, void* d_bytes, uint64_t n, cudaStream_t s) = 0;
  // product-of-pairings check (pairing.cuh): d_g1 / d_g2 = n affine points each; d_f = n * gt_bytes() scratch for the
  // Miller values, d_gt = the reduced pairing of the product, d_flags[0..n) = 1 where P_i is outside the order-r subgroup,
  // d_flags[n] = 1 when the product is one
  virtual size_t gt_bytes() const = 0;
  virtual void pairing_check(const void* d_g1, const void* d_g2, uint32_t n, void* d_f, void* d_gt, uint32_t* d_flags,
                             cudaStream_t s) = 0;
  // n_checks independent checks of `per` pairs each, one thread per pair / per check; d_results[c] = 1 holds, 0 does not,
  // -1 some G1 point of check c is outside the subgroup.  d_f: per * n_checks * gt_bytes(), d_flags: per * n_checks
  virtual void pairing_check_batch(const void* d_g1, const void* d_g2, uint32_t per, uint32_t n_checks, void* d_f,
                                   uint32_t* d_flags, int32_t* d_results, cudaStream_t s) = 0;
  virtual void blob_to_scalars(const void* d_blob, void* d_scalars, uint32_t n, uint32_t* d_err, cudaStream_t s) = 0;
  // KZG opening (EIP-4844 compute_kzg_proof_impl): roots = w^brp(i) table; kzg_open turns the blob's scalars p and
  // the point z (32 big-endian bytes on the device) into the quotient evaluations q and y = p(z) (32 bytes).
  // scratch: (n + 2 * ceil(n / 128) + 2) fr elements + 4 bytes
  virtual void kzg_roots(void* d_roots, uint32_t n, cudaStream_t s) = 0;
  virtual void kzg_open(const void* d_p, const void* d_roots, const void* d_z_be, uint32_t n, void* d_q, void* d_y_bytes,
                        void* d_scratch, uint32_t* d_err, cudaStream_t s) = 0;
  // cell proofs (EIP-7594): a_k table, and the 128 quotients p div (X^m - a_k) from the coefficient form
  virtual void kzg_cell_shifts(void* d_shifts, uint32_t ncells, uint32_t m, uint32_t n, cudaStream_t s) = 0;
  virtual void kzg_cell_quotients(const void* d_coeffs, const void* d_shifts, void* d_q, uint32_t n, uint32_t m,
                                  uint32_t ncells, cudaStream_t s) = 0;
  // --- point helpers
  virtual void to_affine(int group, const void* d_xyzz, void* d_affine, uint32_t count, cudaStream_t s) = 0;
  // affine(sum of count XYZZ points): combine of range-split MSM partials
  // range-split prove: per-slot sums of nparts sets of [5 G1 XYZZ | 1 G2 XYZZ]
  virtual void sum_sets(const void* d_partials, uint32_t nparts, void* d_sums, cudaStream_t s) = 0;
  virtual void sum_partials(int group, const void* d_xyzz, uint32_t count, void* d_affine, cudaStream_t s) = 0;
  // out = sum_i [k_i] P_i  (tiny, single thread): used for the alpha/beta/delta terms of the proof
  // --- debug / parity entry points (element-wise, device pointers)
  virtual void dbg_field_op(int field, int op, const void* a, const void* b, void* out, uint64_t n,
                            cudaStream_t s) = 0;
  virtual void dbg_ec_op(int group, int op, const void* a, const void* b, void* out, uint64_t n,
                         cudaStream_t s) = 0;
  // --- throughput calibration: `iters` dependent Montgomery multiplications per thread (fp)
  virtual void calib_mul(int field, void* d_inout, uint64_t nthreads, int iters, cudaStream_t s) = 0;
  // --- fixed-base batch scalar multiplication: out[i] = affine([k_i] base)   (Setup building block)
  virtual void fixed_base(int group, const void* d_base_affine, const void* d_scalars_mont, uint64_t n,
                          void* d_out_affine, cudaStream_t s) = 0;
};

CurveBackend* backend_bn254();
CurveBackend* backend_bls12_377();
CurveBackend* backend_bls12_381();
CurveBackend* backend_bw6_761();

inline CurveBackend* backend_by_id(int id) {
  switch (id) {
    case 1: return backend_bn254();
    case 2: return backend_bls12_377();
    case 3: return backend_bls12_381();
    case 4: return backend_bw6_761();
    default: return nullptr;
  }
}

}  // namespace b200
