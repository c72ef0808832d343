// Device-resident proving keys / SRS and the proving schedule (see prover.cu).
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/b200_groth16.h"
#include "backend.h"

namespace b200 {

// Pageable host memory (a Go heap slice, a plain numpy array) reaches the GPU through the driver's own single-threaded
// staging at a fraction of the PCIe rate (measured: 11-12 proofs/s instead of 17 for the 537 MB of a voteverifier
// proof).  The stager copies such a source through a ring of page-locked 4 MB chunks with several host threads, each
// chunk followed by its own asynchronous H2D copy, so a caller never has to cudaHostRegister per proof.
struct HostStager {
  static constexpr size_t kChunk = 4u << 20;
  static constexpr int kThreads = 4, kRing = 4;
  uint8_t* pinned = nullptr;
  cudaEvent_t ev[kThreads][kRing] = {};
  int device = 0;
  void init(int dev);
  ~HostStager();
  // enqueue dst[0..bytes) <- src on `s`; src may be pageable, page-locked or (kind == DeviceToDevice) device memory
  void copy(void* d_dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
};

// Workspace + streams for ONE proof in flight.  A key instance owns several slots so that the
// input copies, MSM tails and assembly of one proof overlap the bulk kernels of the next one.
struct PkSlot {
  int device;
  std::mutex mu;
  cudaStream_t st[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t hi = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  DevBuf W, a, b, c, rs, cvals, chal, msm_out, tmp, out_aff;
  MsmWorkspace ws[4];   // [0] Z, [1] wire-indexed G1 sets (+ the shared sort), [2] G2, [3] PoK
  HostStager stager;
  explicit PkSlot(int dev);
  ~PkSlot();
};

constexpr int kSlotsPerDevice = 4;        // proofs in flight per key and GPU (2 for domains >= 2^23, B200_SLOTS overrides)

// HBM budget of ONE key's window tables in bytes (0 = automatic: a share of the device memory free at
// registration).  The registration picks the smallest table stride whose tables fit (msm_plan.h).
void set_pk_table_budget(uint64_t bytes);
uint64_t pk_table_budget();

// One copy of a proving key on one GPU.
struct PkInstance {
  int device;
  // every base set lives as precomputed window tables (MsmBases, table mode): A||delta||alpha,
  // B||delta||beta (G1 and G2), K||delta, Z, the concatenated sigma bases and each commitment basis
  MsmBases tA, tB1, tB2, tK, tZ, tSigma;
  std::vector<std::unique_ptr<MsmBases>> tBasis;
  DevBuf mapA, mapB, mapK, gens;
  NttDomain dom;
  std::vector<std::unique_ptr<PkSlot>> slots;
  std::atomic<uint32_t> next_slot{0};
  int tstride = 1;           // table stride chosen for this key (HBM budget policy)
  uint64_t table_bytes = 0;  // resident window-table bytes of this key on this GPU
  explicit PkInstance(int dev, int nslots);
  // returns a slot with its mutex HELD (first free one, else waits on the next in round-robin order)
  PkSlot& acquire();
};

struct ProvingKeyDev {
  CurveBackend* cb = nullptr;
  uint64_t n = 0, m = 0, nb_public = 0, nA = 0, nB = 0, nK = 0, nZ = 0, total_commit = 0, z_offset = 0;
  int logn = 0;
  std::vector<uint64_t> commit_n;
  std::vector<std::unique_ptr<PkInstance>> inst;
  std::atomic<uint32_t> rr{0};

  static std::unique_ptr<ProvingKeyDev> create(const b200_pk_desc& d, const std::vector<int>& devices);
  PkInstance& pick(int device);
  // d_partials != nullptr: range-split mode - skip the assembly and emit the six raw partial sums
  void prove(const b200_prove_in& in, const b200_proof_out& out, int device, bool inputs_on_device,
             void* d_partials = nullptr);
  void commit(uint32_t i, const b200_slice& values, void* out_affine, int device);
};

struct KzgSrsDev {
  struct Inst {
    int device = 0;
    std::mutex mu;
    cudaStream_t st = nullptr;
    MsmBases tables;
    DevBuf brp, blob, scalars, out, err, roots, quot, scratch, zbuf;
    MsmWorkspace ws;
    // cell proofs: monomial-basis tables, a_k table, 128 quotients, proofs
    MsmBases mono;
    bool have_mono = false;
    NttDomain dom;
    DevBuf shifts, cellq, cell_xyzz, cell_out;
    ~Inst();
  };
  CurveBackend* cb = nullptr;
  uint32_t npoints = 0;
  std::vector<std::unique_ptr<Inst>> inst;
  std::atomic<uint32_t> rr{0};
  static std::unique_ptr<KzgSrsDev> create(const uint8_t* g1_lagrange_compressed, uint32_t npoints,
                                           const std::vector<int>& devices);
  void blob_commit(const uint8_t* blob, uint8_t* commitment48, int device);
  // KZG opening at z (32 big-endian bytes): 48-byte proof and the claimed value y = p(z) (32 big-endian bytes)
  void blob_proof(const uint8_t* blob, const uint8_t* z32, uint8_t* proof48, uint8_t* y32, int device);
  // monomial-basis G1 points [tau^j]_1 (compressed, j < npoints): needed by the cell proofs
  void add_monomial(const uint8_t* g1_monomial_compressed, uint32_t n);
  // EIP-7594 cell proofs: 128 x 48 bytes
  void blob_cell_proofs(const uint8_t* blob, uint8_t* proofs, int device);
  static constexpr uint32_t kCells = 128, kCellSize = 64;
  Inst& pick(int device);
};

}  // namespace b200
