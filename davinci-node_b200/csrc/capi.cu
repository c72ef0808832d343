// C ABI of libb200groth16.so (see include/b200_groth16.h for the contract and the reference
// interfaces each symbol replaces).  Host-side glue only: argument checking, device selection,
// workspace pooling, error capture.  No arithmetic happens on the CPU.
#include "../../include/b200_groth16.h"

#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "backend.h"
#include "msm_plan.h"
#include "prover.h"

using namespace b200;

// ---------------------------------------------------------------------------------- instrumentation
namespace b200 {
namespace {
std::atomic<uint64_t> g_launch_count{0};
std::atomic<bool> g_prof_on{false};
struct ProfRec {
  int tag;
  cudaEvent_t e0, e1;
  cudaStream_t stream;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace
void prof_count_launches(uint64_t n) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
uint64_t prof_launches() { return g_launch_count.load(); }
bool prof_enabled() { return g_prof_on.load(); }
int prof_begin(int tag, cudaStream_t s) {
  if (!g_prof_on.load()) return -1;
  ProfRec r{tag, nullptr, nullptr, s};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  cudaEventRecord(r.e0, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_end(int token, cudaStream_t s) {
  if (token < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if ((size_t)token < g_prof.size()) cudaEventRecord(g_prof[token].e1, s);
}
}  // namespace b200

namespace {

thread_local std::string g_err;

int fail(const std::string& m) {
  g_err = m;
  return 1;
}

template <class Fn>
int guarded(Fn fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    return fail(e.what());
  } catch (...) {
    return fail("unknown error");
  }
}

std::mutex g_mu;
uint32_t g_mask = 0;
bool g_inited = false;

// one workspace per (device, stream); guarded by g_mu for lookup only
std::map<std::pair<int, void*>, std::unique_ptr<MsmWorkspace>> g_ws;

MsmWorkspace& workspace_for(int dev, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto& slot = g_ws[{dev, stream}];
  if (!slot) slot.reset(new MsmWorkspace());
  return *slot;
}

CurveBackend& curve(int id) {
  CurveBackend* b = backend_by_id(id);
  if (!b) throw std::runtime_error("unsupported curve id " + std::to_string(id));
  return *b;
}

void check_group(int g) {
  if (g != 1 && g != 2) throw std::runtime_error("group must be 1 (G1) or 2 (G2)");
}

int current_device() {
  int d = 0;
  B200_CUDA(cudaGetDevice(&d));
  return d;
}

struct DeviceGuard {
  int prev;
  explicit DeviceGuard(int dev) {
    B200_CUDA(cudaGetDevice(&prev));
    if (dev != prev) B200_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

struct ScopedStream {
  cudaStream_t s = nullptr;
  ScopedStream() { B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
  ~ScopedStream() {
    if (s) cudaStreamDestroy(s);
  }
};

// ---- handle registries
struct DomainHandle {
  CurveBackend* cb;
  int device;
  NttDomain dom;
};
std::mutex g_hmu;
uint64_t g_next_handle = 1;
std::map<uint64_t, std::unique_ptr<DomainHandle>> g_domains;
std::map<uint64_t, std::shared_ptr<ProvingKeyDev>> g_pks;
std::map<uint64_t, std::shared_ptr<KzgSrsDev>> g_srs;
struct BasesHandle {
  CurveBackend* cb;
  int device;
  MsmBases bases;
};
std::map<uint64_t, std::shared_ptr<BasesHandle>> g_bases;

std::vector<int> selected_devices() {
  int n = 0;
  B200_CUDA(cudaGetDeviceCount(&n));
  if (n == 0) throw std::runtime_error("no CUDA device visible (this backend has no CPU fallback)");
  uint32_t mask;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    mask = g_inited ? g_mask : 0xffffffffu;
  }
  std::vector<int> v;
  for (int d = 0; d < n && d < 32; d++)
    if (mask & (1u << d)) v.push_back(d);
  if (v.empty()) throw std::runtime_error("device mask selects no visible GPU");
  return v;
}

std::shared_ptr<ProvingKeyDev> find_pk(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_hmu);
  auto it = g_pks.find(h);
  if (it == g_pks.end()) throw std::runtime_error("unknown proving-key handle");
  return it->second;
}

struct ScopedDev {
  void* p = nullptr;
  explicit ScopedDev(size_t bytes) { B200_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); }
  ~ScopedDev() {
    if (p) cudaFree(p);
  }
};

}  // namespace

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
const char* b200_version(void) { return "b200-groth16 0.1 (sm_100a)"; }

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int b200_init(uint32_t device_mask) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    B200_CUDA(cudaGetDeviceCount(&n));
    if (n == 0) throw std::runtime_error("no CUDA device visible (this backend has no CPU fallback)");
    for (int d = 0; d < n; d++) {
      cudaDeviceProp p;
      B200_CUDA(cudaGetDeviceProperties(&p, d));
      if (p.major != 10)
        throw std::runtime_error("device " + std::to_string(d) + " is sm_" + std::to_string(p.major * 10 + p.minor) +
                                 "; this library contains sm_100a code only");
    }
    g_mask = device_mask ? device_mask : ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u));
    g_inited = true;
  });
}

// Page-lock caller memory (a Go slice that lives across proofs: solver output buffers) so that b200_prove's
// host-to-device copies run at full PCIe rate and asynchronously; pageable memory is staged by the driver.
int b200_host_register(void* ptr, uint64_t bytes) {
  return guarded([&] {
    if (!ptr || !bytes) throw std::runtime_error("null argument");
    B200_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  });
}
int b200_host_unregister(void* ptr) {
  return guarded([&] {
    if (!ptr) throw std::runtime_error("null argument");
    B200_CUDA(cudaHostUnregister(ptr));
  });
}

uint64_t b200_fr_bytes(int c) { return backend_by_id(c) ? backend_by_id(c)->fr_bytes() : 0; }
uint64_t b200_fp_bytes(int c) { return backend_by_id(c) ? backend_by_id(c)->fp_bytes() : 0; }
uint64_t b200_affine_bytes(int c, int g) { return backend_by_id(c) ? backend_by_id(c)->affine_bytes(g) : 0; }
uint64_t b200_xyzz_bytes(int c, int g) { return backend_by_id(c) ? backend_by_id(c)->xyzz_bytes(g) : 0; }

int b200_msm_dev(int curve_id, int group, const void* d_points, const void* d_scalars, uint64_t n, void* d_out,
                 int window_bits, void* stream) {
  return guarded([&] {
    check_group(group);
    if (window_bits < 0 || window_bits > 24) throw std::runtime_error("window_bits out of range");
    CurveBackend& cb = curve(curve_id);
    MsmWorkspace& ws = workspace_for(current_device(), stream);
    cb.msm(group, d_points, d_scalars, n, d_out, ws, (cudaStream_t)stream, window_bits, nullptr);
  });
}

int b200_msm_plan(int curve_id, uint64_t n, int window_bits, uint32_t out[5]) {
  return guarded([&] {
    CurveBackend& cb = curve(curve_id);
    MsmPlan st = make_msm_plan(n, cb.fr_bits(), window_bits);
    out[0] = st.c;
    out[1] = st.nwin;
    out[2] = st.nb;
    out[3] = st.task;
    out[4] = st.group;
  });
}

int b200_to_affine_dev(int curve_id, int group, const void* d_xyzz, void* d_affine, uint32_t count, void* stream) {
  return guarded([&] {
    check_group(group);
    curve(curve_id).to_affine(group, d_xyzz, d_affine, count, (cudaStream_t)stream);
  });
}

int b200_msm(int curve_id, int group, const void* points, const void* scalars, uint64_t n, void* out_affine,
             int device) {
  return guarded([&] {
    check_group(group);
    CurveBackend& cb = curve(curve_id);
    if (!out_affine) throw std::runtime_error("out_affine is null");
    if (n && (!points || !scalars)) throw std::runtime_error("null input with n > 0");
    DeviceGuard dg(device);
    ScopedStream st;
    const size_t pb = cb.affine_bytes(group), sb = cb.fr_bytes();
    ScopedDev dp(n * pb), dsc(n * sb), dout(cb.xyzz_bytes(group)), daff(pb);
    if (n) {
      B200_CUDA(cudaMemcpyAsync(dp.p, points, n * pb, cudaMemcpyHostToDevice, st.s));
      B200_CUDA(cudaMemcpyAsync(dsc.p, scalars, n * sb, cudaMemcpyHostToDevice, st.s));
    }
    MsmWorkspace ws;
    cb.msm(group, dp.p, dsc.p, n, dout.p, ws, st.s, 0, nullptr);
    cb.to_affine(group, dout.p, daff.p, 1, st.s);
    B200_CUDA(cudaMemcpyAsync(out_affine, daff.p, pb, cudaMemcpyDeviceToHost, st.s));
    B200_CUDA(cudaStreamSynchronize(st.s));
  });
}

int b200_dbg_field_op_dev(int curve_id, int field, int op, const void* a, const void* b, void* out, uint64_t n,
                          void* stream) {
  return guarded([&] { curve(curve_id).dbg_field_op(field, op, a, b, out, n, (cudaStream_t)stream); });
}

int b200_dbg_ec_op_dev(int curve_id, int group, int op, const void* a, const void* b, void* out, uint64_t n,
                       void* stream) {
  return guarded([&] {
    check_group(group);
    curve(curve_id).dbg_ec_op(group, op, a, b, out, n, (cudaStream_t)stream);
  });
}

int b200_calib_mul_dev(int curve_id, int field, void* d_inout, uint64_t nthreads, int iters, void* stream) {
  return guarded([&] { curve(curve_id).calib_mul(field, d_inout, nthreads, iters, (cudaStream_t)stream); });
}

// ---------------------------------------------------------------------------------- NTT
int b200_domain_create(int curve_id, uint64_t size, const void* generator, const void* coset_gen, uint64_t* out) {
  return guarded([&] {
    CurveBackend& cb = curve(curve_id);
    if (!generator || !coset_gen || !out) throw std::runtime_error("null argument");
    int logn = 0;
    while ((1ull << logn) < size) logn++;
    if ((1ull << logn) != size) throw std::runtime_error("domain size must be a power of two");
    std::unique_ptr<DomainHandle> h(new DomainHandle());
    h->cb = &cb;
    h->device = current_device();
    ScopedStream st;
    ScopedDev gens(2 * cb.fr_bytes());
    B200_CUDA(cudaMemcpyAsync(gens.p, generator, cb.fr_bytes(), cudaMemcpyHostToDevice, st.s));
    B200_CUDA(cudaMemcpyAsync((uint8_t*)gens.p + cb.fr_bytes(), coset_gen, cb.fr_bytes(), cudaMemcpyHostToDevice, st.s));
    cb.domain_init(h->dom, logn, gens.p, (uint8_t*)gens.p + cb.fr_bytes(), st.s);
    B200_CUDA(cudaStreamSynchronize(st.s));
    std::lock_guard<std::mutex> lk(g_hmu);
    *out = g_next_handle++;
    g_domains[*out] = std::move(h);
  });
}

int b200_domain_release(uint64_t h) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_hmu);
    if (!g_domains.erase(h)) throw std::runtime_error("unknown domain handle");
  });
}

static DomainHandle& find_domain(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_hmu);
  auto it = g_domains.find(h);
  if (it == g_domains.end()) throw std::runtime_error("unknown domain handle");
  return *it->second;
}

int b200_ntt_dev(uint64_t domain, void* d_data, int inverse, int decimation, int coset, void* stream) {
  return guarded([&] {
    DomainHandle& h = find_domain(domain);
    h.cb->ntt(h.dom, d_data, inverse != 0, decimation != 0, coset != 0, (cudaStream_t)stream);
  });
}

int b200_compute_h_dev(uint64_t domain, void* d_a, void* d_b, void* d_c, void* stream) {
  return guarded([&] {
    DomainHandle& h = find_domain(domain);
    h.cb->compute_h(h.dom, d_a, d_b, d_c, (cudaStream_t)stream);
  });
}

// ---------------------------------------------------------------------------------- Groth16
int b200_pk_register(const b200_pk_desc* desc, uint64_t* handle_out) {
  return guarded([&] {
    if (!desc || !handle_out) throw std::runtime_error("null argument");
    std::shared_ptr<ProvingKeyDev> pk(ProvingKeyDev::create(*desc, selected_devices()).release());
    std::lock_guard<std::mutex> lk(g_hmu);
    *handle_out = g_next_handle++;
    g_pks[*handle_out] = pk;
  });
}

int b200_pk_release(uint64_t h) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_hmu);
    if (!g_pks.erase(h)) throw std::runtime_error("unknown proving-key handle");
  });
}

int b200_set_pk_table_budget(uint64_t bytes) {
  set_pk_table_budget(bytes);
  return 0;
}

int b200_pk_info(uint64_t h, uint64_t out[6]) {
  return guarded([&] {
    if (!out) throw std::runtime_error("null argument");
    auto pk = find_pk(h);
    PkInstance& I = *pk->inst.at(0);
    out[0] = (uint64_t)I.tA.tstride;   // as clamped to the number of digit windows
    out[1] = I.table_bytes;
    out[2] = I.slots.size();
    out[3] = (uint64_t)I.tA.c;
    out[4] = (uint64_t)I.tZ.c;
    out[5] = pk->inst.size();
  });
}

int b200_commit(uint64_t h, uint32_t i, b200_slice values, void* out, int device) {
  return guarded([&] {
    if (!out) throw std::runtime_error("null output");
    find_pk(h)->commit(i, values, out, device);
  });
}

int b200_prove(uint64_t h, const b200_prove_in* in, const b200_proof_out* out, int device) {
  return guarded([&] {
    if (!in || !out || !out->ar || !out->bs || !out->krs) throw std::runtime_error("null argument");
    find_pk(h)->prove(*in, *out, device, false);
  });
}

int b200_prove_dev(uint64_t h, const b200_prove_in* in, const b200_proof_out* out, int device) {
  return guarded([&] {
    if (!in || !out || !out->ar || !out->bs || !out->krs) throw std::runtime_error("null argument");
    find_pk(h)->prove(*in, *out, device, true);
  });
}

int b200_prove_partial_dev(uint64_t h, const b200_prove_in* in, void* d_partials_out, int device) {
  return guarded([&] {
    if (!in || !d_partials_out) throw std::runtime_error("null argument");
    b200_proof_out none{};
    find_pk(h)->prove(*in, none, device, true, d_partials_out);
  });
}

int b200_pk_coset_evals_dev(uint64_t h, void* d_vec, int device, void* stream) {
  return guarded([&] {
    if (!d_vec) throw std::runtime_error("null argument");
    auto pk = find_pk(h);
    PkInstance& I = pk->pick(device);
    DeviceGuard dg(I.device);
    pk->cb->coset_evals(I.dom, d_vec, (cudaStream_t)stream);
  });
}

int b200_assemble_dev(int curve_id, const void* d_partials, uint32_t nparts, const void* d_r, const void* d_s,
                      int have_pok, void* d_out, void* stream) {
  return guarded([&] {
    if (!d_partials || !nparts || !d_out) throw std::runtime_error("null argument");
    CurveBackend& cb = curve(curve_id);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t x1 = cb.xyzz_bytes(1), x2 = cb.xyzz_bytes(2), g1b = cb.affine_bytes(1), frb = cb.fr_bytes();
    (void)d_r;   // every slice already folded s*Ar_g + r*Bs1_g into its K partial (b200_prove_partial_dev)
    (void)d_s;
    (void)frb;
    ScopedDev sums(5 * x1 + x2), tmp(2 * x1);
    cb.sum_sets(d_partials, nparts, sums.p, s);
    B200_CUDA(cudaMemsetAsync(tmp.p, 0, 2 * x1, s));   // two points at infinity
    uint8_t* sm = (uint8_t*)sums.p;
    uint8_t* o = (uint8_t*)d_out;
    AssembleArgs aa{};
    aa.ar_msm = sm;
    aa.bs1_msm = sm + x1;
    aa.k_msm = sm + 2 * x1;
    aa.z_msm = sm + 3 * x1;
    aa.pok_msm = have_pok ? sm + 4 * x1 : nullptr;
    aa.bs2_msm = sm + 5 * x1;
    aa.rs = nullptr;
    aa.tmp = tmp.p;
    aa.out_ar = o;
    aa.out_krs = o + g1b;
    aa.out_pok = o + 2 * g1b;
    aa.out_bs = o + 3 * g1b;
    cb.assemble(aa, s, 2);
    B200_CUDA(cudaStreamSynchronize(s));   // scratch is scoped to this call
  });
}

// ---------------------------------------------------------------------------------- KZG
int b200_kzg_srs_register(const uint8_t* g1_lagrange, uint32_t npoints, uint64_t* handle_out) {
  return guarded([&] {
    if (!g1_lagrange || !handle_out) throw std::runtime_error("null argument");
    std::shared_ptr<KzgSrsDev> srs(KzgSrsDev::create(g1_lagrange, npoints, selected_devices()).release());
    std::lock_guard<std::mutex> lk(g_hmu);
    *handle_out = g_next_handle++;
    g_srs[*handle_out] = srs;
  });
}

int b200_kzg_srs_release(uint64_t h) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_hmu);
    if (!g_srs.erase(h)) throw std::runtime_error("unknown SRS handle");
  });
}

int b200_blob_commit(uint64_t h, const uint8_t* blob, uint8_t commitment_out[48], int device) {
  return guarded([&] {
    if (!blob || !commitment_out) throw std::runtime_error("null argument");
    std::shared_ptr<KzgSrsDev> srs;
    {
      std::lock_guard<std::mutex> lk(g_hmu);
      auto it = g_srs.find(h);
      if (it == g_srs.end()) throw std::runtime_error("unknown SRS handle");
      srs = it->second;
    }
    srs->blob_commit(blob, commitment_out, device);
  });
}

int b200_kzg_srs_add_monomial(uint64_t h, const uint8_t* g1_monomial, uint32_t npoints) {
  return guarded([&] {
    if (!g1_monomial) throw std::runtime_error("null argument");
    std::shared_ptr<KzgSrsDev> srs;
    {
      std::lock_guard<std::mutex> lk(g_hmu);
      auto it = g_srs.find(h);
      if (it == g_srs.end()) throw std::runtime_error("unknown SRS handle");
      srs = it->second;
    }
    srs->add_monomial(g1_monomial, npoints);
  });
}

int b200_blob_cell_proofs(uint64_t h, const uint8_t* blob, uint8_t* proofs_out, int device) {
  return guarded([&] {
    if (!blob || !proofs_out) throw std::runtime_error("null argument");
    std::shared_ptr<KzgSrsDev> srs;
    {
      std::lock_guard<std::mutex> lk(g_hmu);
      auto it = g_srs.find(h);
      if (it == g_srs.end()) throw std::runtime_error("unknown SRS handle");
      srs = it->second;
    }
    srs->blob_cell_proofs(blob, proofs_out, device);
  });
}

int b200_blob_proof(uint64_t h, const uint8_t* blob, const uint8_t point_be[32], uint8_t proof_out[48],
                    uint8_t claim_out[32], int device) {
  return guarded([&] {
    if (!blob || !point_be || !proof_out || !claim_out) throw std::runtime_error("null argument");
    std::shared_ptr<KzgSrsDev> srs;
    {
      std::lock_guard<std::mutex> lk(g_hmu);
      auto it = g_srs.find(h);
      if (it == g_srs.end()) throw std::runtime_error("unknown SRS handle");
      srs = it->second;
    }
    srs->blob_proof(blob, point_be, proof_out, claim_out, device);
  });
}

// ---------------------------------------------------------------------------------- pairing check
uint64_t b200_gt_bytes(int c) { return backend_by_id(c) ? backend_by_id(c)->gt_bytes() : 0; }

int b200_pairing_check(int curve_id, const void* g1, const void* g2, uint32_t n, int* result_out, void* gt_out, int device) {
  return guarded([&] {
    CurveBackend& cb = curve(curve_id);
    if (!result_out) throw std::runtime_error("result_out is null");
    if (n && (!g1 || !g2)) throw std::runtime_error("null input with n > 0");
    if (n > 64) throw std::runtime_error("pairing check: at most 64 pairs");
    DeviceGuard dg(device);
    ScopedStream st;
    const size_t b1 = cb.affine_bytes(1), b2 = cb.affine_bytes(2), gb = cb.gt_bytes();
    ScopedDev d1(n * b1), d2(n * b2), df(n * gb), dgt(gb), dfl((n + 1) * sizeof(uint32_t));
    if (n) {
      B200_CUDA(cudaMemcpyAsync(d1.p, g1, n * b1, cudaMemcpyHostToDevice, st.s));
      B200_CUDA(cudaMemcpyAsync(d2.p, g2, n * b2, cudaMemcpyHostToDevice, st.s));
    }
    B200_CUDA(cudaMemsetAsync(dfl.p, 0, (n + 1) * sizeof(uint32_t), st.s));
    cb.pairing_check(d1.p, d2.p, n, df.p, dgt.p, (uint32_t*)dfl.p, st.s);
    std::vector<uint32_t> flags(n + 1);
    B200_CUDA(cudaMemcpyAsync(flags.data(), dfl.p, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st.s));
    if (gt_out) B200_CUDA(cudaMemcpyAsync(gt_out, dgt.p, gb, cudaMemcpyDeviceToHost, st.s));
    B200_CUDA(cudaStreamSynchronize(st.s));
    for (uint32_t i = 0; i < n; i++) {
      if (flags[i]) throw std::runtime_error("pairing check: G1 point " + std::to_string(i) + " is not in the order-r subgroup");
    }
    *result_out = flags[n] ? 1 : 0;
  });
}

int b200_pairing_check_batch(int curve_id, const void* g1, const void* g2, uint32_t per, uint32_t n_checks,
                             int32_t* results_out, int device) {
  return guarded([&] {
    CurveBackend& cb = curve(curve_id);
    if (n_checks && !results_out) throw std::runtime_error("results_out is null");
    const uint64_t n = (uint64_t)per * n_checks;
    if (n && (!g1 || !g2)) throw std::runtime_error("null input with pairs > 0");
    if (per > 64 || n > (1u << 20)) throw std::runtime_error("pairing check batch: at most 64 pairs per check and 2^20 pairs");
    if (!n_checks) return;
    DeviceGuard dg(device);
    ScopedStream st;
    const size_t b1 = cb.affine_bytes(1), b2 = cb.affine_bytes(2), gb = cb.gt_bytes();
    ScopedDev d1(n * b1), d2(n * b2), df(n * gb), dfl(n * sizeof(uint32_t)), dres(n_checks * sizeof(int32_t));
    if (n) {
      B200_CUDA(cudaMemcpyAsync(d1.p, g1, n * b1, cudaMemcpyHostToDevice, st.s));
      B200_CUDA(cudaMemcpyAsync(d2.p, g2, n * b2, cudaMemcpyHostToDevice, st.s));
      B200_CUDA(cudaMemsetAsync(dfl.p, 0, n * sizeof(uint32_t), st.s));
    }
    cb.pairing_check_batch(d1.p, d2.p, per, n_checks, df.p, (uint32_t*)dfl.p, (int32_t*)dres.p, st.s);
    B200_CUDA(cudaMemcpyAsync(results_out, dres.p, n_checks * sizeof(int32_t), cudaMemcpyDeviceToHost, st.s));
    B200_CUDA(cudaStreamSynchronize(st.s));
  });
}

// ---------------------------------------------------------------------------------- key artefacts
uint64_t b200_compressed_bytes(int c, int g) { return backend_by_id(c) && (g == 1 || g == 2) ? backend_by_id(c)->compressed_bytes(g) : 0; }

int b200_points_decompress_dev(int curve_id, int group, const void* d_bytes, uint64_t n, void* d_affine_out,
                               uint32_t* d_err_flags, void* stream) {
  return guarded([&] {
    check_group(group);
    if (n && (!d_bytes || !d_affine_out || !d_err_flags)) throw std::runtime_error("null argument");
    curve(curve_id).points_decompress(group, d_bytes, d_affine_out, n, d_err_flags, (cudaStream_t)stream);
  });
}

int b200_points_compress_dev(int curve_id, int group, const void* d_affine, uint64_t n, void* d_bytes_out, void* stream) {
  return guarded([&] {
    check_group(group);
    if (n && (!d_affine || !d_bytes_out)) throw std::runtime_error("null argument");
    curve(curve_id).points_compress(group, d_affine, d_bytes_out, n, (cudaStream_t)stream);
  });
}

// ---------------------------------------------------------------------------------- setup / instrumentation
int b200_fixed_base_dev(int curve_id, int group, const void* d_base_affine, const void* d_scalars, uint64_t n,
                        void* d_out_affine, void* stream) {
  return guarded([&] {
    check_group(group);
    curve(curve_id).fixed_base(group, d_base_affine, d_scalars, n, d_out_affine, (cudaStream_t)stream);
  });
}

int b200_sum_partials_dev(int curve_id, int group, const void* d_xyzz, uint32_t count, void* d_out_affine,
                          void* stream) {
  return guarded([&] {
    check_group(group);
    curve(curve_id).sum_partials(group, d_xyzz, count, d_out_affine, (cudaStream_t)stream);
  });
}

// ---------------------------------------------------------------------------------- base sets (table mode)
int b200_bases_create_dev(int curve_id, int group, const void* d_points, uint64_t n, int window_bits,
                          uint64_t* handle_out, void* stream) {
  return guarded([&] {
    check_group(group);
    if (!handle_out) throw std::runtime_error("null argument");
    std::shared_ptr<BasesHandle> h(new BasesHandle());
    h->cb = &curve(curve_id);
    h->device = current_device();
    h->cb->build_tables(h->bases, group, d_points, n, window_bits, (cudaStream_t)stream);
    B200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    std::lock_guard<std::mutex> lk(g_hmu);
    *handle_out = g_next_handle++;
    g_bases[*handle_out] = h;
  });
}

int b200_bases_release(uint64_t h) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_hmu);
    if (!g_bases.erase(h)) throw std::runtime_error("unknown base-set handle");
  });
}

int b200_msm_bases_dev(uint64_t handle, const void* d_scalars, uint64_t n, const uint32_t* d_index_map,
                       void* d_out_xyzz, void* stream) {
  return guarded([&] {
    std::shared_ptr<BasesHandle> h;
    {
      std::lock_guard<std::mutex> lk(g_hmu);
      auto it = g_bases.find(handle);
      if (it == g_bases.end()) throw std::runtime_error("unknown base-set handle");
      h = it->second;
    }
    if (!d_index_map && n > h->bases.npts) throw std::runtime_error("more scalars than base points");
    MsmWorkspace& ws = workspace_for(h->device, stream);
    h->cb->msm(h->bases.group, nullptr, d_scalars, n, d_out_xyzz, ws, (cudaStream_t)stream, 0, nullptr, d_index_map,
               &h->bases);
  });
}

uint64_t b200_launch_count(void) { return prof_launches(); }

int b200_profile_enable(int on) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_on.store(on != 0);
  });
}

// Synchronises the device, then sums elapsed ms and launch counts per tag (see ProfTag) and clears.
int b200_profile_collect(double ms_out[5], uint64_t count_out[5]) {
  return guarded([&] {
    B200_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int t = 0; t < PROF_TAGS; t++) {
      ms_out[t] = 0;
      count_out[t] = 0;
    }
    for (auto& r : g_prof) {
      float ms = 0;
      if (r.tag < PROF_TAGS && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        ms_out[r.tag] += ms;
        count_out[r.tag]++;
      }
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    g_prof.clear();
  });
}

// Timeline of every instrumented phase since b200_profile_enable(1): 4 doubles per record
// [tag, stream ordinal, start ms, end ms] relative to the earliest start.  Synchronises; does not clear.
int b200_profile_timeline(double* out, uint64_t cap_records, uint64_t* n_out) {
  return guarded([&] {
    B200_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    *n_out = 0;
    if (g_prof.empty()) return;
    std::vector<cudaStream_t> streams;
    std::vector<double> rel(g_prof.size() * 2);
    double lo = 0;
    for (size_t i = 0; i < g_prof.size(); i++) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, g_prof[0].e0, g_prof[i].e0);
      cudaEventElapsedTime(&b, g_prof[0].e0, g_prof[i].e1);
      rel[2 * i] = a;
      rel[2 * i + 1] = b;
      lo = std::min<double>(lo, a);
    }
    for (size_t i = 0; i < g_prof.size() && *n_out < cap_records; i++) {
      size_t k = 0;
      while (k < streams.size() && streams[k] != g_prof[i].stream) k++;
      if (k == streams.size()) streams.push_back(g_prof[i].stream);
      double* o = out + 4 * (*n_out)++;
      o[0] = g_prof[i].tag;
      o[1] = (double)k;
      o[2] = rel[2 * i] - lo;
      o[3] = rel[2 * i + 1] - lo;
    }
  });
}

}  // extern "C"
