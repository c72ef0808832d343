// Groth16 proving orchestration on top of the per-curve kernels: device-resident proving keys,
// the per-proof stream schedule (quotient H + five MSMs + assembly) and the EIP-4844 blob
// commitment.  Host code only sequences kernels and copies; all arithmetic is on the GPU.
//
// Mirrors gnark's `groth16.Prove` data flow behind
// /root/reference/prover/prover_cpu.go:31-58 and prover_gpu.go:24-164.
#include "prover.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace b200 {

namespace {

struct DeviceScope {
  int prev = 0;
  explicit DeviceScope(int dev) {
    B200_CUDA(cudaGetDevice(&prev));
    if (prev != dev) B200_CUDA(cudaSetDevice(dev));
  }
  ~DeviceScope() { cudaSetDevice(prev); }
};

int ilog2_exact(uint64_t n) {
  int l = 0;
  while ((1ull << l) < n) l++;
  if ((1ull << l) != n) throw std::runtime_error("domain size must be a power of two");
  return l;
}

void upload(DevBuf& buf, const void* src, size_t bytes, cudaStream_t s) {
  void* p = buf.get(bytes ? bytes : 16);
  if (bytes) B200_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, s));
}

// every stream of the slot is idle afterwards (the slot's buffers may be reused by the next proof)
void sync_bulk(PkSlot& S) {
  for (auto& s : S.st) B200_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

// ------------------------------------------------------------------------------------ HostStager
void HostStager::init(int dev) {
  device = dev;
  if (pinned) return;
  B200_CUDA(cudaHostAlloc((void**)&pinned, (size_t)kThreads * kRing * kChunk, cudaHostAllocPortable));
  for (auto& row : ev)
    for (auto& e : row) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

HostStager::~HostStager() {
  if (pinned) cudaFreeHost(pinned);
  for (auto& row : ev)
    for (auto& e : row)
      if (e) cudaEventDestroy(e);
}

void HostStager::copy(void* d_dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  if (!bytes) return;
  static const bool enabled = [] {
    const char* e = std::getenv("B200_STAGER");
    return !(e && std::atoi(e) == 0);
  }();
  bool pageable = false;
  if (enabled && kind == cudaMemcpyHostToDevice && bytes >= (8u << 20)) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();   // unknown pointer: treat as page-locked, clear the sticky-free error
  }
  if (!pageable) {
    B200_CUDA(cudaMemcpyAsync(d_dst, src, bytes, kind, s));
    return;
  }
  init(device);
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  std::string errs[kThreads];
  auto worker = [&](int t) {
    try {
      B200_CUDA(cudaSetDevice(device));
      for (size_t c = t, k = 0; c < nchunks; c += kThreads, k++) {
        const int slot = (int)(k % kRing);
        uint8_t* stage = pinned + ((size_t)t * kRing + slot) * kChunk;
        if (k >= (size_t)kRing) B200_CUDA(cudaEventSynchronize(ev[t][slot]));   // its previous DMA has drained
        const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
        std::memcpy(stage, (const uint8_t*)src + off, len);
        B200_CUDA(cudaMemcpyAsync((uint8_t*)d_dst + off, stage, len, cudaMemcpyHostToDevice, s));
        B200_CUDA(cudaEventRecord(ev[t][slot], s));
      }
      // the ring may be reused by the next copy() right away: wait for this thread's last DMAs
      for (int slot = 0; slot < kRing; slot++) B200_CUDA(cudaEventSynchronize(ev[t][slot]));
    } catch (const std::exception& e) {
      errs[t] = e.what();
    }
  };
  std::thread th[kThreads];
  for (int t = 1; t < kThreads; t++) th[t] = std::thread(worker, t);
  worker(0);
  for (int t = 1; t < kThreads; t++) th[t].join();
  for (auto& e : errs)
    if (!e.empty()) throw CudaError(e);
}

// ------------------------------------------------------------------------------------ PkInstance
PkSlot::PkSlot(int dev) : device(dev) {
  stager.device = dev;
  DeviceScope ds(dev);
  // st[0] bulk kernels, st[1] input copies, st[2] / st[3] extra bulk streams (B200_BULK_STREAMS=3);
  // `hi` runs the single-thread proof assembly at the highest priority
  int lo = 0, hi_p = 0;
  B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi_p));
  B200_CUDA(cudaStreamCreateWithPriority(&st[0], cudaStreamNonBlocking, lo));
  B200_CUDA(cudaStreamCreateWithPriority(&st[1], cudaStreamNonBlocking, lo));
  B200_CUDA(cudaStreamCreateWithPriority(&st[2], cudaStreamNonBlocking, lo));
  B200_CUDA(cudaStreamCreateWithPriority(&st[3], cudaStreamNonBlocking, lo));
  B200_CUDA(cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, hi_p));
  for (auto& e : ev) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

PkSlot::~PkSlot() {
  cudaSetDevice(device);
  for (auto& s : st)
    if (s) cudaStreamDestroy(s);
  if (hi) cudaStreamDestroy(hi);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
}

PkInstance::PkInstance(int dev, int nslots) : device(dev) {
  for (int i = 0; i < nslots; i++) slots.emplace_back(new PkSlot(dev));
}

namespace {
std::atomic<uint64_t> g_pk_budget{0};
}
void set_pk_table_budget(uint64_t bytes) { g_pk_budget.store(bytes); }
uint64_t pk_table_budget() { return g_pk_budget.load(); }

PkSlot& PkInstance::acquire() {
  for (auto& s : slots)
    if (s->mu.try_lock()) return *s;
  PkSlot& s = *slots[next_slot.fetch_add(1) % slots.size()];
  s.mu.lock();
  return s;
}

// ------------------------------------------------------------------------------------ registration
std::unique_ptr<ProvingKeyDev> ProvingKeyDev::create(const b200_pk_desc& d, const std::vector<int>& devices) {
  CurveBackend* cb = backend_by_id(d.curve);
  if (!cb) throw std::runtime_error("unsupported curve id " + std::to_string(d.curve));
  std::unique_ptr<ProvingKeyDev> pk(new ProvingKeyDev());
  pk->cb = cb;
  pk->n = d.domain_size;
  pk->logn = ilog2_exact(d.domain_size);
  pk->m = d.nb_wires;
  pk->nb_public = d.nb_public;
  if (pk->nb_public > pk->m) throw std::runtime_error("nb_public > nb_wires");
  if (d.infinity_a.len != pk->m || d.infinity_b.len != pk->m)
    throw std::runtime_error("infinity_a / infinity_b must have nb_wires entries");
  if (!d.generator || !d.coset_gen || !d.g1_alpha || !d.g1_beta || !d.g1_delta || !d.g2_beta || !d.g2_delta)
    throw std::runtime_error("null proving-key element");

  const size_t g1b = cb->affine_bytes(1), g2b = cb->affine_bytes(2), frb = cb->fr_bytes();
  const uint8_t* infA = (const uint8_t*)d.infinity_a.ptr;
  const uint8_t* infB = (const uint8_t*)d.infinity_b.ptr;

  // ---- index maps over the extended wire vector  W_ext = [w_0 .. w_{m-1}, r, s, 1, -rs]
  const uint32_t SKIP = 0xffffffffu;
  // (all three are indexed by the position in W_ext so that one sorting pass can feed A, B and K)
  std::vector<uint32_t> mapA(pk->m + 4), mapB(pk->m + 4), mapK(pk->m + 4, SKIP);
  uint64_t ia = 0, ib = 0;
  for (uint64_t i = 0; i < pk->m; i++) {
    mapA[i] = infA[i] ? SKIP : (uint32_t)ia++;
    mapB[i] = infB[i] ? SKIP : (uint32_t)ib++;
  }
  if (ia != d.g1_A.len) throw std::runtime_error("len(G1.A) != nb_wires - NbInfinityA");
  if (ib != d.g1_B.len || ib != d.g2_B.len) throw std::runtime_error("len(G1.B / G2.B) != nb_wires - NbInfinityB");
  pk->nA = ia;
  pk->nB = ib;
  // tail slots: r, s, 1, -rs
  mapA[pk->m + 0] = (uint32_t)ia;       // r * delta
  mapA[pk->m + 1] = SKIP;
  mapA[pk->m + 2] = (uint32_t)ia + 1;   // 1 * alpha
  mapA[pk->m + 3] = SKIP;
  mapB[pk->m + 0] = SKIP;
  mapB[pk->m + 1] = (uint32_t)ib;       // s * delta
  mapB[pk->m + 2] = (uint32_t)ib + 1;   // 1 * beta
  mapB[pk->m + 3] = SKIP;
  const uint32_t* skip = (const uint32_t*)d.krs_skip.ptr;
  uint64_t nskip = d.krs_skip.len, si = 0, ik = 0;
  const uint64_t npriv = pk->m - pk->nb_public;
  for (uint64_t j = 0; j < npriv; j++) {
    uint64_t wire = pk->nb_public + j;
    while (si < nskip && skip[si] < wire) si++;
    if (si < nskip && skip[si] == wire) mapK[wire] = SKIP;
    else mapK[wire] = (uint32_t)ik++;
  }
  if (ik != d.g1_K.len) throw std::runtime_error("len(G1.K) != private wires - skipped wires");
  pk->nK = ik;
  mapK[pk->m + 3] = (uint32_t)ik;       // (-rs) * delta
  pk->nZ = d.g1_Z.len;
  pk->z_offset = d.z_offset;
  if (pk->z_offset + pk->nZ > pk->n) throw std::runtime_error("z_offset + len(G1.Z) > domain size");

  // ---- commitment keys: concatenated sigma bases, per-commitment bases
  pk->commit_n.resize(d.nb_commitments);
  uint64_t total_sigma = 0;
  for (uint32_t i = 0; i < d.nb_commitments; i++) {
    if (d.commit_basis[i].len != d.commit_basis_exp_sigma[i].len)
      throw std::runtime_error("commitment key basis / basisExpSigma length mismatch");
    pk->commit_n[i] = d.commit_basis[i].len;
    total_sigma += pk->commit_n[i];
  }
  pk->total_commit = total_sigma;

  int nslots = pk->n >= (1ull << 23) ? 2 : kSlotsPerDevice;
  if (const char* e = std::getenv("B200_SLOTS")) nslots = std::max(1, std::min(8, std::atoi(e)));
  for (int dev : devices) {
    DeviceScope ds(dev);
    std::unique_ptr<PkInstance> in(new PkInstance(dev, nslots));
    cudaStream_t s = in->slots[0]->st[0];
    // ---- HBM budget policy: the smallest table stride whose window tables (plus the transient scratch of the
    //      largest table build) fit the key's budget.  Budget: b200_set_pk_table_budget / B200_PK_BUDGET_GB, else
    //      55% of the memory free right now minus this key's proof workspaces.
    const uint64_t wire_pts = std::max({d.g1_A.len + 2, d.g1_B.len + 2, d.g1_K.len + 1});
    // Window of the wire-indexed keys (they share one digit / sort pass per proof, hence one width).
    // The cost model assumes dense scalars; a solved witness is sparse (40% zeros, 20% ones, 25% below
    // 2^64), so the wire-indexed sets fill ~3 digits per scalar and their bucket reduction weighs far more than the
    // model thinks.  Measured on the 2^22 voteverifier shape, 4 proofs in flight: c = 20 / 19 / 18 / 17 / 16 -> 17.3 /
    // 17.5 / 18.3 / 17.9 / 14.5 proofs/s (and 8.1 -> 7.7 proofs/s for a uniform-random wire vector at c = 18): two
    // bits below the model, but never so few buckets that the one-thread-per-bucket accumulation runs out of threads.
    // A range-split slice (its Z set covers at most half the domain) is proved one at a time, nothing overlaps its
    // MSMs, and the narrower window starves the G2 accumulation: slices keep the model's window (measured on the
    // BW6-761 2^22 aggregator shape: N = 2 / 4 slices 135.8 / 80.9 ms with the model's window, 172.5 / 98.2 ms two
    // bits below it).
    const bool slice_key = d.g1_Z.len * 2 <= d.domain_size;
    auto wire_window = [&](int ts) {
      int cw = cb->table_window(wire_pts, ts);
      if (!slice_key && cw >= 18) cw = std::max(17, cw - 2);
      if (const char* e = std::getenv("B200_WIRE_WINDOW")) {   // tuning knob
        if (std::atoi(e) >= 4 && std::atoi(e) <= 22) cw = std::atoi(e);
      }
      return cw;
    };
    // G2 re-uses B1's digit / sort pass (same wire map), so it shares the wire window by default.  B200_G2_WINDOW=model
    // (or a width) gives it its own window and its own sort pass: a single proof run alone finishes sooner (2^22
    // voteverifier: 67 -> 58.4 ms, the G2 accumulation no longer starves: ncu 34% -> 82% multiplier pipe) but with
    // four proofs in flight the extra sort costs throughput (18.24 -> 17.50 proofs/s), so it is opt-in.
    auto g2_window = [&](int ts, int cw) {
      if (const char* e = std::getenv("B200_G2_WINDOW")) {
        if (!std::strcmp(e, "model")) return cb->table_window(d.g2_B.len + 2, ts);
        if (std::atoi(e) >= 4 && std::atoi(e) <= 22) return std::atoi(e);
      }
      return cw;
    };
    auto tables_bytes = [&](int ts, uint64_t* peak) {
      const int bits = cb->fr_bits();
      auto set_bytes = [&](uint64_t npts, size_t pb, int c) {
        return (uint64_t)msm_ntables(msm_nwin(bits, c), ts) * std::max<uint64_t>(npts, 1) * pb;
      };
      const int cw = wire_window(ts);
      const int cz = cb->table_window(std::max<uint64_t>(d.g1_Z.len, 1), ts);
      uint64_t parts[6] = {set_bytes(d.g1_A.len + 2, g1b, cw), set_bytes(d.g1_B.len + 2, g1b, cw),
                           set_bytes(d.g2_B.len + 2, g2b, g2_window(ts, cw)),
                           set_bytes(d.g1_K.len + 1, g1b, cw),
                           set_bytes(d.g1_Z.len, g1b, cz),
                           2 * set_bytes(total_sigma, g1b, cb->table_window(std::max<uint64_t>(total_sigma, 1), ts))};
      uint64_t sum = 0, big = 0;
      for (uint64_t v : parts) {
        sum += v;
        big = std::max(big, v);
      }
      *peak = sum + big + big / 2;   // staging copy + 1.5x scratch of the set being built
      return sum;
    };
    {
      uint64_t budget = pk_table_budget();
      if (const char* e = std::getenv("B200_PK_BUDGET_GB")) budget = (uint64_t)(std::atof(e) * 1e9);
      if (!budget) {
        size_t free_b = 0, total_b = 0;
        B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t slot_est = (uint64_t)nslots * (pk->m + 3 * pk->n) * frb * 3;   // inputs + sort / bucket workspaces
        budget = free_b > slot_est ? (uint64_t)((free_b - slot_est) * 0.55) : 0;
      }
      int ts = 1;
      if (const char* e = std::getenv("B200_TABLE_STRIDE")) {
        ts = std::max(1, std::atoi(e));
      } else {
        uint64_t peak = 0;
        while (ts < 32 && (tables_bytes(ts, &peak), peak) > budget) ts++;
      }
      in->tstride = ts;
    }
    const int ts = in->tstride;
    // upload (points || extra0 || extra1) to a staging buffer, build its window tables, drop the staging copy
    DevBuf stage;
    auto put_tables = [&](MsmBases& t, int group, const b200_slice& sl, size_t pb, const void* e0, const void* e1,
                          int window_bits = 0) {
      uint64_t cnt = sl.len + (e0 ? 1 : 0) + (e1 ? 1 : 0);
      uint8_t* p = (uint8_t*)stage.get(std::max<uint64_t>(cnt, 1) * pb);
      if (sl.len) B200_CUDA(cudaMemcpyAsync(p, sl.ptr, sl.len * pb, cudaMemcpyHostToDevice, s));
      if (e0) B200_CUDA(cudaMemcpyAsync(p + sl.len * pb, e0, pb, cudaMemcpyHostToDevice, s));
      if (e1) B200_CUDA(cudaMemcpyAsync(p + (sl.len + 1) * pb, e1, pb, cudaMemcpyHostToDevice, s));
      cb->build_tables(t, group, p, cnt, window_bits, s, ts);
      in->table_bytes += (uint64_t)t.ntab * std::max<uint64_t>(cnt, 1) * pb;
      B200_CUDA(cudaStreamSynchronize(s));   // staging buffer is reused by the next base set
    };
    const int cw = wire_window(ts);
    put_tables(in->tA, 1, d.g1_A, g1b, d.g1_delta, d.g1_alpha, cw);
    put_tables(in->tB1, 1, d.g1_B, g1b, d.g1_delta, d.g1_beta, cw);
    const int cw2 = g2_window(ts, cw);
    put_tables(in->tB2, 2, d.g2_B, g2b, d.g2_delta, d.g2_beta, cw2);
    put_tables(in->tK, 1, d.g1_K, g1b, d.g1_delta, nullptr, cw);
    put_tables(in->tZ, 1, d.g1_Z, g1b, nullptr, nullptr);
    upload(in->mapA, mapA.data(), mapA.size() * 4, s);
    upload(in->mapB, mapB.data(), mapB.size() * 4, s);
    upload(in->mapK, mapK.data(), mapK.size() * 4, s);
    in->tBasis.resize(d.nb_commitments);
    {
      // concatenated sigma bases -> one table set
      uint8_t* sg = (uint8_t*)stage.get(std::max<uint64_t>(total_sigma, 1) * g1b);
      uint64_t off = 0;
      for (uint32_t i = 0; i < d.nb_commitments; i++) {
        if (pk->commit_n[i])
          B200_CUDA(cudaMemcpyAsync(sg + off * g1b, d.commit_basis_exp_sigma[i].ptr, pk->commit_n[i] * g1b,
                                    cudaMemcpyHostToDevice, s));
        off += pk->commit_n[i];
      }
      cb->build_tables(in->tSigma, 1, sg, total_sigma, 0, s, ts);
      in->table_bytes += (uint64_t)in->tSigma.ntab * std::max<uint64_t>(total_sigma, 1) * g1b;
      B200_CUDA(cudaStreamSynchronize(s));
    }
    for (uint32_t i = 0; i < d.nb_commitments; i++) {
      in->tBasis[i].reset(new MsmBases());
      put_tables(*in->tBasis[i], 1, d.commit_basis[i], g1b, nullptr, nullptr);
    }
    // ---- domain tables
    uint8_t* gens = (uint8_t*)in->gens.get(2 * frb);
    B200_CUDA(cudaMemcpyAsync(gens, d.generator, frb, cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaMemcpyAsync(gens + frb, d.coset_gen, frb, cudaMemcpyHostToDevice, s));
    cb->domain_init(in->dom, pk->logn, gens, gens + frb, s);
    // ---- per-proof workspace
    for (auto& sl : in->slots) {
      sl->W.get((pk->m + 4) * frb);
      sl->a.get(pk->n * frb);
      sl->b.get(pk->n * frb);
      sl->c.get(pk->n * frb);
      sl->rs.get(2 * frb);
      sl->cvals.get(std::max<uint64_t>(total_sigma, 1) * frb);
      sl->chal.get(frb);
      sl->msm_out.get(6 * cb->xyzz_bytes(2));
      sl->tmp.get(2 * cb->xyzz_bytes(1));
      sl->out_aff.get(3 * g1b + g2b);
    }
    B200_CUDA(cudaStreamSynchronize(s));
    pk->inst.push_back(std::move(in));
  }
  return pk;
}

PkInstance& ProvingKeyDev::pick(int device) {
  if (device >= 0) {
    for (auto& in : inst)
      if (in->device == device) return *in;
    throw std::runtime_error("proving key is not resident on device " + std::to_string(device));
  }
  uint32_t k = rr.fetch_add(1) % (uint32_t)inst.size();
  return *inst[k];
}

// ------------------------------------------------------------------------------------ prove
void ProvingKeyDev::prove(const b200_prove_in& in, const b200_proof_out& out, int device, bool inputs_on_device,
                          void* d_partials) {
  PkInstance& I = pick(device);
  PkSlot& S = I.acquire();
  std::lock_guard<std::mutex> lk(S.mu, std::adopt_lock);
  DeviceScope ds(I.device);
  const size_t frb = cb->fr_bytes(), g1b = cb->affine_bytes(1), g2b = cb->affine_bytes(2);
  const size_t x1 = cb->xyzz_bytes(1), x2 = cb->xyzz_bytes(2);
  if (in.wires.len != m) throw std::runtime_error("wires: expected nb_wires elements");
  if (in.a.len > n || in.b.len != in.a.len || in.c.len != in.a.len)
    throw std::runtime_error("a/b/c: need equal lengths <= domain size");
  if (in.abc_form > 1) throw std::runtime_error("abc_form must be 0 or 1");
  if (in.abc_form == 1 && in.a.len != n) throw std::runtime_error("abc_form = 1 needs domain_size coset evaluations");
  if (in.nb_commitments != commit_n.size()) throw std::runtime_error("nb_commitments mismatch with proving key");
  if (!in.r || !in.s) throw std::runtime_error("r / s missing");
  if ((m && !in.wires.ptr) || (in.a.len && (!in.a.ptr || !in.b.ptr || !in.c.ptr)))
    throw std::runtime_error("wires / a / b / c: null pointer");
  // every argument is validated before the first copy is enqueued
  if (!commit_n.empty()) {
    if (!in.priv_committed) throw std::runtime_error("priv_committed is null but the key has commitments");
    for (size_t i = 0; i < commit_n.size(); i++) {
      if (in.priv_committed[i].len != commit_n[i])
        throw std::runtime_error("private committed values: length mismatch with commitment key");
      if (commit_n[i] && !in.priv_committed[i].ptr) throw std::runtime_error("private committed values: null pointer");
    }
    if (commit_n.size() > 1 && !in.fold_challenge)
      throw std::runtime_error("fold_challenge required with more than one commitment");
  }
  const cudaMemcpyKind kind = inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  // Streams.  sc: input copies.  Bulk (throughput-bound) kernels - sorts, bucket accumulation, NTT passes - go to
  // ONE stream per proof by default (sb == sw == sg): two different accumulation kernels sharing an SM evict each
  // other's instruction-cache-sized loop bodies, so they are better run back to back.  B200_BULK_STREAMS=3 restores
  // three concurrent bulk streams.  Every MSM's latency-bound tail runs on its workspace's high-priority stream
  // and is only joined by the assembly.
  static const bool three_bulk = [] {
    const char* e = std::getenv("B200_BULK_STREAMS");
    return e && std::atoi(e) == 3;
  }();
  cudaStream_t sc = S.st[1], sb = S.st[0];
  cudaStream_t sw = three_bulk ? S.st[2] : sb;    // wire-indexed G1 MSMs
  cudaStream_t sg = three_bulk ? S.st[3] : sb;    // G2 MSM
  uint8_t* W = (uint8_t*)S.W.p;
  uint8_t* a = (uint8_t*)S.a.p;
  uint8_t* b = (uint8_t*)S.b.p;
  uint8_t* c = (uint8_t*)S.c.p;

  // ---- inputs (sc): wire vector first so the wire-indexed MSMs can start while a, b, c are still arriving
  int tok_in = prof_begin(PROF_INPUTS, sc);
  S.stager.copy(W, in.wires.ptr, m * frb, kind, sc);
  uint8_t* rs_in = (uint8_t*)S.rs.p;
  B200_CUDA(cudaMemcpyAsync(rs_in, in.r, frb, kind, sc));
  B200_CUDA(cudaMemcpyAsync(rs_in + frb, in.s, frb, kind, sc));
  cb->prep_rs(rs_in, rs_in + frb, W + m * frb, sc);
  prof_end(tok_in, sc);
  B200_CUDA(cudaEventRecord(S.ev[0], sc));
  tok_in = prof_begin(PROF_INPUTS, sc);
  const uint64_t nc = in.a.len;
  if (nc < n) {
    B200_CUDA(cudaMemsetAsync(a + nc * frb, 0, (n - nc) * frb, sc));
    B200_CUDA(cudaMemsetAsync(b + nc * frb, 0, (n - nc) * frb, sc));
    B200_CUDA(cudaMemsetAsync(c + nc * frb, 0, (n - nc) * frb, sc));
  }
  if (nc) {
    S.stager.copy(a, in.a.ptr, nc * frb, kind, sc);
    S.stager.copy(b, in.b.ptr, nc * frb, kind, sc);
    S.stager.copy(c, in.c.ptr, nc * frb, kind, sc);
  }
  bool have_pok = total_commit > 0 || !commit_n.empty();
  uint8_t* cv = (uint8_t*)S.cvals.p;
  if (have_pok) {
    uint64_t off = 0;
    for (size_t i = 0; i < commit_n.size(); i++) {
      if (commit_n[i]) B200_CUDA(cudaMemcpyAsync(cv + off * frb, in.priv_committed[i].ptr, commit_n[i] * frb, kind, sc));
      if (i == 1) B200_CUDA(cudaMemcpyAsync(S.chal.p, in.fold_challenge, frb, kind, sc));
      off += commit_n[i];
    }
  }
  prof_end(tok_in, sc);
  B200_CUDA(cudaEventRecord(S.ev[1], sc));

  uint8_t* mo = (uint8_t*)S.msm_out.p;   // [ar, bs1, k, z, pok] G1 then bs2 G2
  uint8_t *o_ar = mo, *o_z = mo + 3 * x1, *o_pok = mo + 4 * x1, *o_bs2 = mo + 5 * x1;

  // ---- wire-indexed MSMs: ONE digit/sort pass over W_ext feeds the A, B and K keys; Ar, Bs1 and the wire part
  //      of Krs are reduced together in G1, Bs (G2) reduces the B set of the same sorted data
  B200_CUDA(cudaStreamWaitEvent(sw, S.ev[0], 0));
  {
    const MsmBases* sets[3] = {&I.tA, &I.tB1, &I.tK};
    const uint32_t* maps[3] = {(const uint32_t*)I.mapA.p, (const uint32_t*)I.mapB.p, (const uint32_t*)I.mapK.p};
    MsmSorted so;
    cb->msm_sort(W, m + 4, sets, maps, 3, S.ws[1], sw, so);
    if (sg != sw) {
      B200_CUDA(cudaEventRecord(S.ev[2], sw));
      B200_CUDA(cudaStreamWaitEvent(sg, S.ev[2], 0));
    }
    cb->msm_reduce(1, so, 0, 3, sets, o_ar, S.ws[1], sw, false);   // -> o_ar, o_bs1, o_k (contiguous)
    if (I.tB2.c == I.tB1.c && I.tB2.tstride == I.tB1.tstride) {
      const MsmBases* g2set[1] = {&I.tB2};
      cb->msm_reduce(2, so, 1, 1, g2set, o_bs2, S.ws[2], sg, false);
    } else {
      cb->msm(2, nullptr, W, m + 4, o_bs2, S.ws[2], sg, 0, nullptr, (const uint32_t*)I.mapB.p, &I.tB2, false);
    }
  }
  // s*Ar and r*Bs1 (two single-thread scalar multiplications, 1.4 ms on BLS12-377, 10 ms on BW6-761) only need the
  // wire-indexed G1 sums: they run on the high-priority stream as soon as that tail is done, long before the quotient MSM
  cudaStream_t sh = S.hi;
  AssembleArgs aa{};
  aa.ar_msm = o_ar;
  aa.bs1_msm = mo + x1;
  aa.rs = W + m * frb;
  aa.tmp = S.tmp.p;
  // (range-split slices do the same with their partial sums: sum_g s*Ar_g = s*Ar, so the multiplications never sit
  // on the critical path after the gather)
  aa.k_msm = mo + 2 * x1;
  {
    S.ws[1].wait_tail(sh);
    const int tok_mul = prof_begin(PROF_ASSEMBLE, sh);
    cb->assemble(aa, sh, d_partials ? (1 | 4) : 1);
    prof_end(tok_mul, sh);
  }

  // ---- proof of knowledge, quotient, Z MSM.  The (small) PoK MSM goes first: its latency-bound tail then hides behind
  //      the quotient and the Z accumulation instead of ending the proof
  B200_CUDA(cudaStreamWaitEvent(sb, S.ev[1], 0));
  if (have_pok) {
    // segment i of the committed values is scaled by challenge^i: scale every later segment once per step
    uint64_t off = 0;
    for (size_t i = 0; i < commit_n.size(); i++) {
      if (i >= 1) cb->scale_vec(cv + off * frb, S.chal.p, total_commit - off, sb);
      off += commit_n[i];
    }
    cb->msm(1, nullptr, cv, total_commit, o_pok, S.ws[3], sb, 0, nullptr, nullptr, &I.tSigma, false);
  }
  if (in.abc_form == 1) cb->compute_h_tail(I.dom, a, b, c, sb);
  else cb->compute_h(I.dom, a, b, c, sb);
  cb->msm(1, nullptr, a + z_offset * frb, nZ, o_z, S.ws[0], sb, 0, nullptr, nullptr, &I.tZ, false);
  // ---- join every tail on the high-priority stream
  S.ws[0].wait_tail(sh);
  S.ws[1].wait_tail(sh);
  S.ws[2].wait_tail(sh);
  if (have_pok) S.ws[3].wait_tail(sh);
  if (d_partials) {
    // range-split mode: hand the partial sums back (layout of msm_out: 5 G1 XYZZ then 1 G2 XYZZ); the K slot
    // already holds K_g + s*Ar_g + r*Bs1_g
    if (!have_pok) B200_CUDA(cudaMemsetAsync(o_pok, 0, x1, sh));
    B200_CUDA(cudaMemcpyAsync(d_partials, mo, 5 * x1 + x2, cudaMemcpyDeviceToDevice, sh));
    B200_CUDA(cudaStreamSynchronize(sh));
    sync_bulk(S);
    return;
  }

  uint8_t* oa = (uint8_t*)S.out_aff.p;   // ar, krs, pok (G1), bs (G2)
  aa.bs2_msm = o_bs2;
  aa.k_msm = mo + 2 * x1;
  aa.z_msm = o_z;
  aa.pok_msm = have_pok ? o_pok : nullptr;
  aa.out_ar = oa;
  aa.out_krs = oa + g1b;
  aa.out_pok = oa + 2 * g1b;
  aa.out_bs = oa + 3 * g1b;
  // the assembly is two single-thread scalar multiplications: it runs on the high-priority stream so it is
  // not queued behind the accumulation blocks of the other proofs in flight
  const int tok_asm = prof_begin(PROF_ASSEMBLE, S.hi);
  cb->assemble(aa, S.hi, 2);
  prof_end(tok_asm, S.hi);
  const cudaMemcpyKind back = inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  B200_CUDA(cudaMemcpyAsync(out.ar, oa, g1b, back, S.hi));
  B200_CUDA(cudaMemcpyAsync(out.krs, oa + g1b, g1b, back, S.hi));
  if (have_pok && out.pok) B200_CUDA(cudaMemcpyAsync(out.pok, oa + 2 * g1b, g1b, back, S.hi));
  B200_CUDA(cudaMemcpyAsync(out.bs, oa + 3 * g1b, g2b, back, S.hi));
  B200_CUDA(cudaStreamSynchronize(S.hi));
  sync_bulk(S);
}

// commitment i = sum_j values[j] * Basis_i[j]   (called from the solver hint, synchronous)
void ProvingKeyDev::commit(uint32_t i, const b200_slice& values, void* out_affine, int device) {
  if (i >= commit_n.size()) throw std::runtime_error("commitment index out of range");
  if (values.len != commit_n[i]) throw std::runtime_error("commit: values length mismatch with Basis");
  PkInstance& I = pick(device);
  PkSlot& S = I.acquire();
  std::lock_guard<std::mutex> lk(S.mu, std::adopt_lock);
  DeviceScope ds(I.device);
  const size_t frb = cb->fr_bytes(), g1b = cb->affine_bytes(1);
  cudaStream_t s0 = S.st[0];
  uint8_t* cv = (uint8_t*)S.cvals.p;
  if (values.len) B200_CUDA(cudaMemcpyAsync(cv, values.ptr, values.len * frb, cudaMemcpyHostToDevice, s0));
  uint8_t* mo = (uint8_t*)S.msm_out.p;
  cb->msm(1, nullptr, cv, values.len, mo, S.ws[0], s0, 0, nullptr, nullptr, I.tBasis[i].get());
  cb->to_affine(1, mo, S.out_aff.p, 1, s0);
  B200_CUDA(cudaMemcpyAsync(out_affine, S.out_aff.p, g1b, cudaMemcpyDeviceToHost, s0));
  B200_CUDA(cudaStreamSynchronize(s0));
}

// ------------------------------------------------------------------------------------ KZG
std::unique_ptr<KzgSrsDev> KzgSrsDev::create(const uint8_t* g1_lagrange_compressed, uint32_t npoints,
                                             const std::vector<int>& devices) {
  CurveBackend* cb = backend_by_id(3);
  if (npoints == 0 || (npoints & (npoints - 1))) throw std::runtime_error("SRS size must be a power of two");
  std::unique_ptr<KzgSrsDev> srs(new KzgSrsDev());
  srs->cb = cb;
  srs->npoints = npoints;
  int logn = ilog2_exact(npoints);
  std::vector<uint32_t> brp(npoints);
  for (uint32_t i = 0; i < npoints; i++) {
    uint32_t r = 0;
    for (int b = 0; b < logn; b++) r |= ((i >> b) & 1u) << (logn - 1 - b);
    brp[i] = r;
  }
  for (int dev : devices) {
    DeviceScope ds(dev);
    std::unique_ptr<KzgSrsDev::Inst> in(new KzgSrsDev::Inst());
    in->device = dev;
    B200_CUDA(cudaStreamCreateWithFlags(&in->st, cudaStreamNonBlocking));
    const size_t g1b = cb->affine_bytes(1);
    DevBuf raw;
    upload(raw, g1_lagrange_compressed, (size_t)npoints * 48, in->st);
    DevBuf pts;
    pts.get((size_t)npoints * g1b);
    uint32_t* err = (uint32_t*)in->err.get(4);
    B200_CUDA(cudaMemsetAsync(err, 0, 4, in->st));
    cb->g1_decompress(raw.p, pts.p, npoints, err, in->st);
    cb->build_tables(in->tables, 1, pts.p, npoints, 0, in->st);
    upload(in->brp, brp.data(), (size_t)npoints * 4, in->st);
    in->blob.get((size_t)npoints * 32);
    in->scalars.get((size_t)npoints * cb->fr_bytes());
    in->out.get(cb->xyzz_bytes(1) + g1b + 64);
    const size_t frb = cb->fr_bytes();
    cb->kzg_roots(in->roots.get((size_t)npoints * frb), npoints, in->st);
    in->quot.get((size_t)npoints * frb);
    in->scratch.get(((size_t)npoints + 2 * ((npoints + 127) / 128) + 2) * frb + 16);
    in->zbuf.get(64);
    uint32_t herr = 0;
    B200_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, in->st));
    B200_CUDA(cudaStreamSynchronize(in->st));
    if (herr) throw std::runtime_error("SRS contains an invalid compressed G1 point (flags " + std::to_string(herr) + ")");
    srs->inst.push_back(std::move(in));
  }
  return srs;
}

KzgSrsDev::Inst::~Inst() {
  cudaSetDevice(device);
  if (st) cudaStreamDestroy(st);
}

void KzgSrsDev::add_monomial(const uint8_t* g1_monomial_compressed, uint32_t n) {
  if (n != npoints) throw std::runtime_error("monomial SRS must have as many points as the Lagrange SRS");
  if (npoints < 2 * kCellSize) throw std::runtime_error("SRS too small for cell proofs");
  for (auto& in : inst) {
    std::lock_guard<std::mutex> lk(in->mu);
    DeviceScope ds(in->device);
    const size_t g1b = cb->affine_bytes(1), frb = cb->fr_bytes();
    DevBuf raw, pts;
    upload(raw, g1_monomial_compressed, (size_t)n * 48, in->st);
    pts.get((size_t)n * g1b);
    uint32_t* err = (uint32_t*)in->err.p;
    B200_CUDA(cudaMemsetAsync(err, 0, 4, in->st));
    cb->g1_decompress(raw.p, pts.p, n, err, in->st);
    cb->build_tables(in->mono, 1, pts.p, n, 0, in->st);
    uint32_t herr = 0;
    B200_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, in->st));
    B200_CUDA(cudaStreamSynchronize(in->st));
    if (herr) throw std::runtime_error("monomial SRS contains an invalid compressed G1 point");
    // evaluation domain of the blob (omega = roots[n/2] since roots[i] = omega^brp(i); the coset generator is unused)
    const uint8_t* roots = (const uint8_t*)in->roots.p;
    cb->domain_init(in->dom, ilog2_exact(npoints), roots + (size_t)(npoints / 2) * frb, roots + frb, in->st);
    cb->kzg_cell_shifts(in->shifts.get((size_t)kCells * frb), kCells, kCellSize, npoints, in->st);
    in->cellq.get((size_t)kCells * npoints * frb);
    in->cell_xyzz.get((size_t)kCells * cb->xyzz_bytes(1));
    in->cell_out.get((size_t)kCells * (g1b + 48));
    B200_CUDA(cudaStreamSynchronize(in->st));
    in->have_mono = true;
  }
}

void KzgSrsDev::blob_cell_proofs(const uint8_t* blob, uint8_t* proofs, int device) {
  Inst* I = &pick(device);
  std::lock_guard<std::mutex> lk(I->mu);
  if (!I->have_mono) throw std::runtime_error("cell proofs need the monomial SRS (b200_kzg_srs_add_monomial)");
  DeviceScope ds(I->device);
  const size_t g1b = cb->affine_bytes(1), x1 = cb->xyzz_bytes(1), frb = cb->fr_bytes();
  uint32_t* err = (uint32_t*)I->err.p;
  B200_CUDA(cudaMemsetAsync(err, 0, 4, I->st));
  B200_CUDA(cudaMemcpyAsync(I->blob.p, blob, (size_t)npoints * 32, cudaMemcpyHostToDevice, I->st));
  cb->blob_to_scalars(I->blob.p, I->scalars.p, npoints, err, I->st);
  // evaluations in bit-reversed order -> coefficients: inverse DIT transform (bit-reversed in, natural out, 1/n)
  cb->ntt(I->dom, I->scalars.p, true, true, false, I->st);
  cb->kzg_cell_quotients(I->scalars.p, I->shifts.p, I->cellq.p, npoints, kCellSize, kCells, I->st);
  // the 128 quotients against the monomial basis: ONE batched MSM (one sort, one accumulate / reduce pass)
  uint8_t* xyzz = (uint8_t*)I->cell_xyzz.p;
  cb->msm_batch(1, I->mono, I->cellq.p, npoints, kCells, nullptr, xyzz, I->ws, I->st);
  uint8_t* o = (uint8_t*)I->cell_out.p;
  cb->to_affine(1, xyzz, o, kCells, I->st);
  cb->g1_compress(o, o + (size_t)kCells * g1b, kCells, I->st);
  uint32_t herr = 0;
  B200_CUDA(cudaMemcpyAsync(proofs, o + (size_t)kCells * g1b, (size_t)kCells * 48, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaStreamSynchronize(I->st));
  if (herr) throw std::runtime_error("blob contains a non-canonical field element (>= BLS12-381 r)");
}

KzgSrsDev::Inst& KzgSrsDev::pick(int device) {
  if (device >= 0) {
    for (auto& in : inst)
      if (in->device == device) return *in;
    throw std::runtime_error("SRS is not resident on device " + std::to_string(device));
  }
  return *inst[rr.fetch_add(1) % inst.size()];
}

void KzgSrsDev::blob_proof(const uint8_t* blob, const uint8_t* z32, uint8_t* proof48, uint8_t* y32, int device) {
  Inst* I = &pick(device);
  std::lock_guard<std::mutex> lk(I->mu);
  DeviceScope ds(I->device);
  const size_t g1b = cb->affine_bytes(1), x1 = cb->xyzz_bytes(1);
  uint32_t* err = (uint32_t*)I->err.p;
  uint8_t* zb = (uint8_t*)I->zbuf.p;   // [z (32) | y (32)]
  B200_CUDA(cudaMemsetAsync(err, 0, 4, I->st));
  B200_CUDA(cudaMemcpyAsync(I->blob.p, blob, (size_t)npoints * 32, cudaMemcpyHostToDevice, I->st));
  B200_CUDA(cudaMemcpyAsync(zb, z32, 32, cudaMemcpyHostToDevice, I->st));
  cb->blob_to_scalars(I->blob.p, I->scalars.p, npoints, err, I->st);
  cb->kzg_open(I->scalars.p, I->roots.p, zb, npoints, I->quot.p, zb + 32, I->scratch.p, err, I->st);
  uint8_t* o = (uint8_t*)I->out.p;
  cb->msm(1, nullptr, I->quot.p, npoints, o, I->ws, I->st, 0, nullptr, (const uint32_t*)I->brp.p, &I->tables);
  cb->to_affine(1, o, o + x1, 1, I->st);
  cb->g1_compress(o + x1, o + x1 + g1b, 1, I->st);
  uint32_t herr = 0;
  B200_CUDA(cudaMemcpyAsync(proof48, o + x1 + g1b, 48, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaMemcpyAsync(y32, zb + 32, 32, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaStreamSynchronize(I->st));
  if (herr & 8u) throw std::runtime_error("evaluation point is not a canonical field element (>= BLS12-381 r)");
  if (herr) throw std::runtime_error("blob contains a non-canonical field element (>= BLS12-381 r)");
}

void KzgSrsDev::blob_commit(const uint8_t* blob, uint8_t* commitment48, int device) {
  Inst* I = &pick(device);
  std::lock_guard<std::mutex> lk(I->mu);
  DeviceScope ds(I->device);
  const size_t g1b = cb->affine_bytes(1), x1 = cb->xyzz_bytes(1);
  uint32_t* err = (uint32_t*)I->err.p;
  B200_CUDA(cudaMemsetAsync(err, 0, 4, I->st));
  B200_CUDA(cudaMemcpyAsync(I->blob.p, blob, (size_t)npoints * 32, cudaMemcpyHostToDevice, I->st));
  cb->blob_to_scalars(I->blob.p, I->scalars.p, npoints, err, I->st);
  uint8_t* o = (uint8_t*)I->out.p;
  cb->msm(1, nullptr, I->scalars.p, npoints, o, I->ws, I->st, 0, nullptr, (const uint32_t*)I->brp.p, &I->tables);
  cb->to_affine(1, o, o + x1, 1, I->st);
  cb->g1_compress(o + x1, o + x1 + g1b, 1, I->st);
  uint32_t herr = 0;
  B200_CUDA(cudaMemcpyAsync(commitment48, o + x1 + g1b, 48, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, I->st));
  B200_CUDA(cudaStreamSynchronize(I->st));
  if (herr) throw std::runtime_error("blob contains a non-canonical field element (>= BLS12-381 r)");
}

}  // namespace b200
