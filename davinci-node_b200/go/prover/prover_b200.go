//go:build b200

// Package prover - B200 backend, common part.  Drop the files of this directory into
// /root/reference/prover/ and build with `-tags b200` (and add `&& !b200` to the build constraints of
// prover_cpu.go / prover_gpu.go, see INTEGRATION.md).  Every exported name and signature of
// prover_cpu.go:19-64 and prover_gpu.go:66-164 is kept, so circuits/artifacts.go:480,579,591 and the
// sequencer / worker flows compile and run unchanged.
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain there).  The files are pure marshalling: every
// arithmetic step is behind the C ABI in include/b200_groth16.h, which the Python mirror
// (davinci-node_b200/prover.py) drives with the same buffers and which the GPU parity tests cover;
// tests/test_go_shim.py checks that every C.b200_* call names a symbol of the header with the header's
// arity.  The per-curve files prover_b200_<curve>.go are generated from curve.go.tmpl by gen_curves.py
// (the reference's own callGPUProver switches per curve the same way, prover_gpu.go:24-61).
package prover

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../lib -lb200groth16 -Wl,-rpath,${SRCDIR}/../lib
#include <stdlib.h>
#include "b200_groth16.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc"
	fr_bls12377 "github.com/consensys/gnark-crypto/ecc/bls12-377/fr"
	fr_bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381/fr"
	fr_bn254 "github.com/consensys/gnark-crypto/ecc/bn254/fr"
	fr_bw6761 "github.com/consensys/gnark-crypto/ecc/bw6-761/fr"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/groth16"
	groth16_bls12377 "github.com/consensys/gnark/backend/groth16/bls12-377"
	groth16_bls12381 "github.com/consensys/gnark/backend/groth16/bls12-381"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	groth16_bw6761 "github.com/consensys/gnark/backend/groth16/bw6-761"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs_bls12377 "github.com/consensys/gnark/constraint/bls12-377"
	cs_bls12381 "github.com/consensys/gnark/constraint/bls12-381"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
	cs_bw6761 "github.com/consensys/gnark/constraint/bw6-761"
	"github.com/consensys/gnark/frontend"
)

// b200Call runs one C-ABI call and, on failure, fetches its message.  The library keeps the message in a
// thread-local: the goroutine is locked to its OS thread for the pair of calls so it cannot migrate
// between the failing call and b200_last_error (ADVICE r1).
func b200Call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if f() != 0 {
		return fmt.Errorf("b200: %s", C.GoString(C.b200_last_error()))
	}
	return nil
}

var (
	initOnce sync.Once
	initErr  error
)

func b200Init() error {
	initOnce.Do(func() {
		initErr = b200Call(func() C.int { return C.b200_init(0) })
	})
	return initErr
}

func b200Slice[T any](v []T) C.b200_slice {
	if len(v) == 0 {
		return C.b200_slice{}
	}
	return C.b200_slice{ptr: unsafe.Pointer(&v[0]), len: C.uint64_t(len(v))}
}

func boolsToBytes(b []bool) []byte {
	// []bool is one byte per element (0/1) in Go; reinterpret without copying
	if len(b) == 0 {
		return nil
	}
	return unsafe.Slice((*byte)(unsafe.Pointer(&b[0])), len(b))
}

// krsSkipList: the wires left out of the K MSM (every PrivateCommitted wire and every CommitmentIndex wire), ascending.
func krsSkipList(info constraint.Groth16Commitments) []uint32 {
	var skip []uint32
	for i := range info {
		for _, w := range info[i].PrivateCommitted {
			skip = append(skip, uint32(w))
		}
		skip = append(skip, uint32(info[i].CommitmentIndex))
	}
	for i := 1; i < len(skip); i++ { // insertion sort: the list is nearly sorted
		for j := i; j > 0 && skip[j-1] > skip[j]; j-- {
			skip[j-1], skip[j] = skip[j], skip[j-1]
		}
	}
	return skip
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
// nothing in this package calls it - a GPU error is returned to the caller, never retried on the CPU.
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
		cpk, ok := pk.(*groth16_bn254.ProvingKey)
		if !ok {
			return nil, fmt.Errorf("proving key type mismatch for BN254: expected *groth16_bn254.ProvingKey, got %T", pk)
		}
		r1cs, ok := ccs.(*cs_bn254.R1CS)
		if !ok {
			return nil, fmt.Errorf("constraint system type mismatch for BN254: got %T", ccs)
		}
		return proveBN254(r1cs, cpk, w, opts...)

	case ecc.BLS12_377:
		cpk, ok := pk.(*groth16_bls12377.ProvingKey)
		if !ok {
			return nil, fmt.Errorf("proving key type mismatch for BLS12_377: expected *groth16_bls12377.ProvingKey, got %T", pk)
		}
		r1cs, ok := ccs.(*cs_bls12377.R1CS)
		if !ok {
			return nil, fmt.Errorf("constraint system type mismatch for BLS12_377: got %T", ccs)
		}
		// (the reference sets PinToGPU for this hot circuit, prover_gpu.go:40-41: here every key stays resident)
		return proveBLS12377(r1cs, cpk, w, opts...)

	case ecc.BLS12_381:
		cpk, ok := pk.(*groth16_bls12381.ProvingKey)
		if !ok {
			return nil, fmt.Errorf("proving key type mismatch for BLS12_381: expected *groth16_bls12381.ProvingKey, got %T", pk)
		}
		r1cs, ok := ccs.(*cs_bls12381.R1CS)
		if !ok {
			return nil, fmt.Errorf("constraint system type mismatch for BLS12_381: got %T", ccs)
		}
		return proveBLS12381(r1cs, cpk, w, opts...)

	case ecc.BW6_761:
		cpk, ok := pk.(*groth16_bw6761.ProvingKey)
		if !ok {
			return nil, fmt.Errorf("proving key type mismatch for BW6_761: expected *groth16_bw6761.ProvingKey, got %T", pk)
		}
		r1cs, ok := ccs.(*cs_bw6761.R1CS)
		if !ok {
			return nil, fmt.Errorf("constraint system type mismatch for BW6_761: got %T", ccs)
		}
		return proveBW6761(r1cs, cpk, w, opts...)

	default:
		return nil, fmt.Errorf("B200 proving not supported for curve %s", curveID)
	}
}

// VerifyB200 is groth16.Verify (circuits/artifacts.go:604-613) with its MultiExp and pairing checks on the GPU.  Same
// arguments, same error cases; opt-in - the reference's own groth16.Verify call keeps working unchanged.
func VerifyB200(proof groth16.Proof, vk groth16.VerifyingKey, publicWitness witness.Witness, opts ...backend.VerifierOption) error {
	switch p := proof.(type) {
	case *groth16_bn254.Proof:
		k, ok := vk.(*groth16_bn254.VerifyingKey)
		w, ok2 := publicWitness.Vector().(fr_bn254.Vector)
		if !ok || !ok2 {
			return fmt.Errorf("verifying key / witness type mismatch for BN254: got %T, %T", vk, publicWitness.Vector())
		}
		return verifyBN254(p, k, w, opts...)
	case *groth16_bls12377.Proof:
		k, ok := vk.(*groth16_bls12377.VerifyingKey)
		w, ok2 := publicWitness.Vector().(fr_bls12377.Vector)
		if !ok || !ok2 {
			return fmt.Errorf("verifying key / witness type mismatch for BLS12_377: got %T, %T", vk, publicWitness.Vector())
		}
		return verifyBLS12377(p, k, w, opts...)
	case *groth16_bls12381.Proof:
		k, ok := vk.(*groth16_bls12381.VerifyingKey)
		w, ok2 := publicWitness.Vector().(fr_bls12381.Vector)
		if !ok || !ok2 {
			return fmt.Errorf("verifying key / witness type mismatch for BLS12_381: got %T, %T", vk, publicWitness.Vector())
		}
		return verifyBLS12381(p, k, w, opts...)
	case *groth16_bw6761.Proof:
		k, ok := vk.(*groth16_bw6761.VerifyingKey)
		w, ok2 := publicWitness.Vector().(fr_bw6761.Vector)
		if !ok || !ok2 {
			return fmt.Errorf("verifying key / witness type mismatch for BW6_761: got %T, %T", vk, publicWitness.Vector())
		}
		return verifyBW6761(p, k, w, opts...)
	default:
		return fmt.Errorf("B200 verification not supported for proof type %T", proof)
	}
}
