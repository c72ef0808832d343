//go:build b200

// Package prover - B200 backend.  Drop this file into /root/reference/prover/ and build with
// `-tags b200` (and add `&& !b200` to the build constraints of prover_cpu.go / prover_gpu.go, see
// INTEGRATION.md).  It keeps every exported name and signature of prover_cpu.go:19-64 and
// prover_gpu.go:66-164, so circuits/artifacts.go:480,579,591 and the sequencer / worker flows compile
// and run unchanged.
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain there).  It is pure marshalling: every
// arithmetic step is behind the C ABI in include/b200_groth16.h, which the Python mirror
// (davinci-node_b200/prover.py) drives with the same buffers and which the GPU parity tests cover.
// The BN254 instantiation is spelled out; BLS12-377 / BW6-761 are the same code over the sibling
// gnark packages (the reference's own callGPUProver switches per curve the same way,
// prover_gpu.go:24-61).
package prover

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../lib -lb200groth16 -Wl,-rpath,${SRCDIR}/../lib
#include <stdlib.h>
#include "b200_groth16.h"
*/
import "C"

import (
	"crypto/rand"
	"fmt"
	"math/big"
	"runtime"
	"sort"
	"sync"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc"
	curve "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/hash_to_field"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/groth16"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/constraint/solver"
	"github.com/consensys/gnark/frontend"
	fcs "github.com/consensys/gnark/frontend/cs"
)

func lastErr() error { return fmt.Errorf("b200: %s", C.GoString(C.b200_last_error())) }

// ---- pinned-randomness hook (test only; SURVEY.md 8b).  nil = crypto/rand like gnark.
var randomnessHook func() (r, s fr.Element)

// SetRandomnessForTest pins the prover randomness; exported names of the package are untouched.
func SetRandomnessForTest(f func() (r, s fr.Element)) { randomnessHook = f }

func sampleRS() (r, s fr.Element, err error) {
	if randomnessHook != nil {
		r, s = randomnessHook()
		return
	}
	if _, err = r.SetRandom(); err != nil {
		return
	}
	_, err = s.SetRandom()
	_ = rand.Reader
	return
}

// ---- device-resident proving keys, keyed by the *ProvingKey the caller keeps for the process
// lifetime (circuits/artifacts.go:518-523)
var (
	pkMu      sync.Mutex
	pkHandles = map[*groth16_bn254.ProvingKey]C.uint64_t{}
	initOnce  sync.Once
	initErr   error
)

func slice[T any](v []T) C.b200_slice {
	if len(v) == 0 {
		return C.b200_slice{}
	}
	return C.b200_slice{ptr: unsafe.Pointer(&v[0]), len: C.uint64_t(len(v))}
}

func registerBN254(pk *groth16_bn254.ProvingKey, r1cs *cs.R1CS) (C.uint64_t, error) {
	initOnce.Do(func() {
		if C.b200_init(0) != 0 {
			initErr = lastErr()
		}
	})
	if initErr != nil {
		return 0, initErr
	}
	pkMu.Lock()
	defer pkMu.Unlock()
	if h, ok := pkHandles[pk]; ok {
		return h, nil
	}
	info := r1cs.CommitmentInfo.(constraint.Groth16Commitments)
	var skip []uint32
	for i := range info {
		for _, w := range info[i].PrivateCommitted {
			skip = append(skip, uint32(w))
		}
		skip = append(skip, uint32(info[i].CommitmentIndex))
	}
	sort.Slice(skip, func(a, b int) bool { return skip[a] < skip[b] })
	infA := boolsToBytes(pk.InfinityA)
	infB := boolsToBytes(pk.InfinityB)

	// cgo rule: a C struct passed to C may not contain Go pointers to Go pointers; the slice arrays
	// are therefore allocated in C memory, the point/scalar payloads are pinned for the call.
	var pin runtime.Pinner
	defer pin.Unpin()
	p := func(ptr unsafe.Pointer) unsafe.Pointer { pin.Pin(ptr); return ptr }
	k := len(pk.CommitmentKeys)
	basis := (*[1 << 20]C.b200_slice)(C.malloc(C.size_t(max(k, 1)) * C.size_t(unsafe.Sizeof(C.b200_slice{}))))
	sigma := (*[1 << 20]C.b200_slice)(C.malloc(C.size_t(max(k, 1)) * C.size_t(unsafe.Sizeof(C.b200_slice{}))))
	defer C.free(unsafe.Pointer(basis))
	defer C.free(unsafe.Pointer(sigma))
	for i := range pk.CommitmentKeys {
		basis[i] = slice(pk.CommitmentKeys[i].Basis)
		sigma[i] = slice(pk.CommitmentKeys[i].BasisExpSigma)
		p(basis[i].ptr)
		p(sigma[i].ptr)
	}
	d := C.b200_pk_desc{
		curve:       C.B200_BN254,
		domain_size: C.uint64_t(pk.Domain.Cardinality),
		generator:   p(unsafe.Pointer(&pk.Domain.Generator)),
		coset_gen:   p(unsafe.Pointer(&pk.Domain.FrMultiplicativeGen)),
		g1_alpha:    p(unsafe.Pointer(&pk.G1.Alpha)), g1_beta: p(unsafe.Pointer(&pk.G1.Beta)), g1_delta: p(unsafe.Pointer(&pk.G1.Delta)),
		g1_A:        slice(pk.G1.A), g1_B: slice(pk.G1.B), g1_Z: slice(pk.G1.Z), g1_K: slice(pk.G1.K),
		g2_beta:     p(unsafe.Pointer(&pk.G2.Beta)), g2_delta: p(unsafe.Pointer(&pk.G2.Delta)),
		g2_B:        slice(pk.G2.B),
		infinity_a:  slice(infA), infinity_b: slice(infB),
		nb_wires:    C.uint64_t(len(pk.InfinityA)),
		nb_public:   C.uint64_t(r1cs.GetNbPublicVariables()),
		krs_skip:    slice(skip),
		nb_commitments:         C.uint32_t(k),
		commit_basis:           &basis[0],
		commit_basis_exp_sigma: &sigma[0],
	}
	for _, s := range []C.b200_slice{d.g1_A, d.g1_B, d.g1_Z, d.g1_K, d.g2_B, d.infinity_a, d.infinity_b, d.krs_skip} {
		if s.ptr != nil {
			p(s.ptr)
		}
	}
	var h C.uint64_t
	if C.b200_pk_register(&d, &h) != 0 {
		return 0, lastErr()
	}
	pkHandles[pk] = h
	return h, nil
}

func boolsToBytes(b []bool) []byte {
	// []bool is one byte per element (0/1) in Go; reinterpret without copying
	if len(b) == 0 {
		return nil
	}
	return unsafe.Slice((*byte)(unsafe.Pointer(&b[0])), len(b))
}

// proveBN254 is gnark's groth16.Prove (SURVEY.md A.1) with computeH, the five MultiExps, the
// Pedersen commitment / proof of knowledge and the final assembly replaced by the C ABI.
func proveBN254(r1cs *cs.R1CS, pk *groth16_bn254.ProvingKey, fullWitness witness.Witness, opts ...backend.ProverOption) (*groth16_bn254.Proof, error) {
	opt, err := backend.NewProverConfig(opts...)
	if err != nil {
		return nil, fmt.Errorf("new prover config: %w", err)
	}
	if opt.HashToFieldFn == nil {
		opt.HashToFieldFn = hash_to_field.New([]byte(constraint.CommitmentDst))
	}
	h, err := registerBN254(pk, r1cs)
	if err != nil {
		return nil, err
	}
	commitmentInfo := r1cs.CommitmentInfo.(constraint.Groth16Commitments)
	proof := &groth16_bn254.Proof{Commitments: make([]curve.G1Affine, len(commitmentInfo))}
	privateCommittedValues := make([][]fr.Element, len(commitmentInfo))
	solverOpts := opt.SolverOpts[:len(opt.SolverOpts):len(opt.SolverOpts)]

	// BSB22 hint override: identical to gnark's, except Commit() is one GPU MSM (b200_commit)
	for i := range commitmentInfo {
		solverOpts = append(solverOpts, solver.OverrideHint(commitmentInfo[i].HintID, func(i int) solver.Hint {
			return func(_ *big.Int, in []*big.Int, out []*big.Int) error {
				privateCommittedValues[i] = make([]fr.Element, len(commitmentInfo[i].PrivateCommitted))
				hashed := in[:len(commitmentInfo[i].PublicAndCommitmentCommitted)]
				committed := in[+len(hashed):]
				for j, inJ := range committed {
					privateCommittedValues[i][j].SetBigInt(inJ)
				}
				vals := privateCommittedValues[i]
				var pin runtime.Pinner
				defer pin.Unpin()
				sl := slice(vals)
				if sl.ptr != nil {
					pin.Pin(sl.ptr)
				}
				if C.b200_commit(h, C.uint32_t(i), sl, unsafe.Pointer(&proof.Commitments[i]), -1) != 0 {
					return lastErr()
				}
				opt.HashToFieldFn.Write(constraint.SerializeCommitment(proof.Commitments[i].Marshal(), hashed, (fr.Bits-1)/8+1))
				hashBts := opt.HashToFieldFn.Sum(nil)
				opt.HashToFieldFn.Reset()
				nbBuf := fr.Bytes
				if opt.HashToFieldFn.Size() < fr.Bytes {
					nbBuf = opt.HashToFieldFn.Size()
				}
				var res fr.Element
				res.SetBytes(hashBts[:nbBuf])
				res.BigInt(out[0])
				return nil
			}
		}(i)))
	}

	_solution, err := r1cs.Solve(fullWitness, solverOpts...) // host, unchanged gnark solver
	if err != nil {
		return nil, err
	}
	solution := _solution.(*cs.R1CSSolution)

	var foldChallenge fr.Element
	if len(commitmentInfo) > 1 {
		commitmentsSerialized := make([]byte, fr.Bytes*len(commitmentInfo))
		for i := range commitmentInfo {
			copy(commitmentsSerialized[fr.Bytes*i:], solution.W[commitmentInfo[i].CommitmentIndex].Marshal())
		}
		ch, err := fr.Hash(commitmentsSerialized, []byte("G16-BSB22"), 1)
		if err != nil {
			return nil, err
		}
		foldChallenge = ch[0]
	}
	r, s, err := sampleRS()
	if err != nil {
		return nil, err
	}

	var pin runtime.Pinner
	defer pin.Unpin()
	k := len(commitmentInfo)
	pcs := (*[1 << 20]C.b200_slice)(C.malloc(C.size_t(max(k, 1)) * C.size_t(unsafe.Sizeof(C.b200_slice{}))))
	defer C.free(unsafe.Pointer(pcs))
	for i := range privateCommittedValues {
		pcs[i] = slice(privateCommittedValues[i])
		if pcs[i].ptr != nil {
			pin.Pin(pcs[i].ptr)
		}
	}
	in := C.b200_prove_in{
		wires: slice([]fr.Element(solution.W)), a: slice([]fr.Element(solution.A)),
		b: slice([]fr.Element(solution.B)), c: slice([]fr.Element(solution.C)),
		r: unsafe.Pointer(&r), s: unsafe.Pointer(&s),
		nb_commitments: C.uint32_t(k), priv_committed: &pcs[0],
		fold_challenge: unsafe.Pointer(&foldChallenge),
	}
	for _, sl := range []C.b200_slice{in.wires, in.a, in.b, in.c} {
		if sl.ptr != nil {
			pin.Pin(sl.ptr)
		}
	}
	pin.Pin(in.r)
	pin.Pin(in.s)
	pin.Pin(in.fold_challenge)
	out := C.b200_proof_out{
		ar: unsafe.Pointer(&proof.Ar), bs: unsafe.Pointer(&proof.Bs),
		krs: unsafe.Pointer(&proof.Krs), pok: unsafe.Pointer(&proof.CommitmentPok),
	}
	pin.Pin(out.ar)
	if C.b200_prove(h, &in, &out, -1) != 0 {
		return nil, lastErr() // no CPU fallback (north_star); callers already treat err as "batch failed"
	}
	return proof, nil
}

// ---------------------------------------------------------------- exported surface (unchanged)

// Prove runs groth16.Prove on the B200 backend (prover_cpu.go:19 / prover_gpu.go:66).
func Prove(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, assignment frontend.Circuit, opts ...backend.ProverOption) (groth16.Proof, error) {
	return prover(curveID, ccs, pk, assignment, opts...)
}

func defaultProver(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, assignment frontend.Circuit, opts ...backend.ProverOption) (groth16.Proof, error) {
	return GPUProver(curveID, ccs, pk, assignment, opts...)
}

// CPUProver keeps gnark's CPU path available under its reference name (prover_cpu.go:31-38);
// nothing in this file calls it - a GPU error is returned to the caller, never retried on the CPU.
func CPUProver(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, assignment frontend.Circuit, opts ...backend.ProverOption) (groth16.Proof, error) {
	w, err := frontend.NewWitness(assignment, curveID.ScalarField())
	if err != nil {
		return nil, fmt.Errorf("failed to create witness: %w", err)
	}
	return groth16.Prove(ccs, pk, w, opts...)
}

// GPUProver: prover_gpu.go:86-96.
func GPUProver(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, assignment frontend.Circuit, opts ...backend.ProverOption) (groth16.Proof, error) {
	w, err := frontend.NewWitness(assignment, curveID.ScalarField())
	if err != nil {
		return nil, fmt.Errorf("failed to create witness: %w", err)
	}
	return GPUProverWithWitness(curveID, ccs, pk, w, opts...)
}

// ProveWithWitness: prover_cpu.go:49 / prover_gpu.go:111.
func ProveWithWitness(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, w witness.Witness, opts ...backend.ProverOption) (groth16.Proof, error) {
	return GPUProverWithWitness(curveID, ccs, pk, w, opts...)
}

// CPUProverWithWitness: prover_cpu.go:53-58.
func CPUProverWithWitness(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, w witness.Witness, opts ...backend.ProverOption) (groth16.Proof, error) {
	return groth16.Prove(ccs, pk, w, opts...)
}

// GPUProverWithWitness: prover_gpu.go:121-131 - curve dispatch with the same pk type assertions as
// callGPUProver (prover_gpu.go:24-61), on PLAIN gnark key types (UseGPUProver may stay false for
// artifact loading, circuits/artifacts.go:624-646).
func GPUProverWithWitness(curveID ecc.ID, ccs constraint.ConstraintSystem, pk groth16.ProvingKey, w witness.Witness, opts ...backend.ProverOption) (groth16.Proof, error) {
	switch curveID {
	case ecc.BN254:
		bnPk, ok := pk.(*groth16_bn254.ProvingKey)
		if !ok {
			return nil, fmt.Errorf("proving key type mismatch for BN254: expected *groth16_bn254.ProvingKey, got %T", pk)
		}
		r1cs, ok := ccs.(*cs.R1CS)
		if !ok {
			return nil, fmt.Errorf("constraint system type mismatch for BN254: got %T", ccs)
		}
		return proveBN254(r1cs, bnPk, w, opts...)
	// case ecc.BLS12_377: proveBLS12377(...)   - same body over ecc/bls12-377, B200_BLS12_377
	// case ecc.BW6_761:   proveBW6761(...)     - same body over ecc/bw6-761,  B200_BW6_761
	default:
		return nil, fmt.Errorf("B200 proving not supported for curve %s", curveID)
	}
}

var _ = fcs.ErrInputNotSet
