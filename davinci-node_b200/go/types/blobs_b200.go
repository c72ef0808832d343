//go:build b200

// B200 replacements of the four KZG methods of types.Blob (/root/reference/types/blobs.go:90-134).  To integrate:
// move the bodies of ComputeCommitment / ComputeCellProofs / ComputeBlobProof / ComputeProof of types/blobs.go into a
// file tagged `//go:build !b200` and add this file; every other method of blobs.go (ComputeCommitmentAndProof,
// ComputeCommitmentAndCellProofs, the sidecar builders) calls these four and stays as it is.
//
// The EIP-4844 ceremony (config.KZGTrustedSetup, /root/reference/config/kzg_setup.go:5-8: 4096 G1 Lagrange points,
// 65 G2 points, 4096 G1 monomial points, hex, one per line) is uploaded once; each call is then one or 128 BLS12-381
// MSMs on the GPU (b200_blob_commit / b200_blob_cell_proofs / b200_blob_proof).  NOT COMPILED IN THE BUILD CONTAINER
// (no Go toolchain): tests/test_go_shim.py checks the C calls against include/b200_groth16.h.
package types

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../lib -lb200groth16 -Wl,-rpath,${SRCDIR}/../lib
#include "b200_groth16.h"
*/
import "C"

import (
	"bytes"
	"crypto/sha256"
	"encoding/hex"
	"fmt"
	"math/big"
	"runtime"
	"strconv"
	"sync"
	"unsafe"

	"github.com/vocdoni/davinci-node/config"
)

var blsModulus, _ = new(big.Int).SetString("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001", 16)

func b200Call(f func() C.int) error {
	runtime.LockOSThread() // the library's error message is thread-local
	defer runtime.UnlockOSThread()
	if f() != 0 {
		return fmt.Errorf("b200: %s", C.GoString(C.b200_last_error()))
	}
	return nil
}

var (
	srsOnce   sync.Once
	srsHandle C.uint64_t
	srsErr    error
)

// parseTrustedSetup splits config.KZGTrustedSetup into the compressed G1 Lagrange and G1 monomial blocks.
func parseTrustedSetup(raw []byte) (lagrange, monomial []byte, err error) {
	lines := bytes.Fields(raw)
	if len(lines) < 2 {
		return nil, nil, fmt.Errorf("trusted setup: too short")
	}
	n1, err := strconv.Atoi(string(lines[0]))
	if err != nil {
		return nil, nil, fmt.Errorf("trusted setup: %w", err)
	}
	n2, err := strconv.Atoi(string(lines[1]))
	if err != nil {
		return nil, nil, fmt.Errorf("trusted setup: %w", err)
	}
	if len(lines) != 2+n1+n2+n1 {
		return nil, nil, fmt.Errorf("trusted setup: expected %d entries, got %d", 2+n1+n2+n1, len(lines))
	}
	decode := func(from int) ([]byte, error) {
		out := make([]byte, 0, n1*48)
		for _, l := range lines[from : from+n1] {
			b, err := hex.DecodeString(string(l))
			if err != nil || len(b) != 48 {
				return nil, fmt.Errorf("trusted setup: bad G1 point %q", l)
			}
			out = append(out, b...)
		}
		return out, nil
	}
	if lagrange, err = decode(2); err != nil {
		return nil, nil, err
	}
	monomial, err = decode(2 + n1 + n2)
	return lagrange, monomial, err
}

func srs() (C.uint64_t, error) {
	srsOnce.Do(func() {
		lag, mono, err := parseTrustedSetup(config.KZGTrustedSetup)
		if err != nil {
			srsErr = err
			return
		}
		if srsErr = b200Call(func() C.int { return C.b200_init(0) }); srsErr != nil {
			return
		}
		n := C.uint32_t(len(lag) / 48)
		if srsErr = b200Call(func() C.int {
			return C.b200_kzg_srs_register((*C.uint8_t)(unsafe.Pointer(&lag[0])), n, &srsHandle)
		}); srsErr != nil {
			return
		}
		srsErr = b200Call(func() C.int {
			return C.b200_kzg_srs_add_monomial(srsHandle, (*C.uint8_t)(unsafe.Pointer(&mono[0])), n)
		})
	})
	return srsHandle, srsErr
}

// ComputeCommitment creates a small commitment out of a data blob (types/blobs.go:90-96).
func (b *Blob) ComputeCommitment() (KZGCommitment, error) {
	h, err := srs()
	if err != nil {
		return KZGCommitment{}, err
	}
	var out KZGCommitment
	err = b200Call(func() C.int {
		return C.b200_blob_commit(h, (*C.uint8_t)(unsafe.Pointer(&b[0])), (*C.uint8_t)(unsafe.Pointer(&out[0])), -1)
	})
	return out, err
}

// ComputeCellProofs returns the 128 EIP-7594 cell proofs of the blob (types/blobs.go:99-105).
func (b *Blob) ComputeCellProofs() ([]KZGProof, error) {
	h, err := srs()
	if err != nil {
		return nil, err
	}
	proofs := make([]KZGProof, CellProofsPerBlob)
	err = b200Call(func() C.int {
		return C.b200_blob_cell_proofs(h, (*C.uint8_t)(unsafe.Pointer(&b[0])), (*C.uint8_t)(unsafe.Pointer(&proofs[0])), -1)
	})
	if err != nil {
		return nil, err
	}
	return proofs, nil
}

// ComputeProof computes the KZG proof at the given point for the polynomial represented by the blob
// (types/blobs.go:123-134).
func (b *Blob) ComputeProof(point *big.Int) (proof KZGProof, claim *big.Int, err error) {
	var z [32]byte
	if point.Sign() < 0 || point.BitLen() > len(z)*8 {
		return KZGProof{}, nil, fmt.Errorf("point does not fit in %d bytes", len(z))
	}
	point.FillBytes(z[:])
	h, err := srs()
	if err != nil {
		return KZGProof{}, nil, err
	}
	var y [32]byte
	err = b200Call(func() C.int {
		return C.b200_blob_proof(h, (*C.uint8_t)(unsafe.Pointer(&b[0])), (*C.uint8_t)(unsafe.Pointer(&z[0])),
			(*C.uint8_t)(unsafe.Pointer(&proof[0])), (*C.uint8_t)(unsafe.Pointer(&y[0])), -1)
	})
	if err != nil {
		return KZGProof{}, nil, err
	}
	return proof, new(big.Int).SetBytes(y[:]), nil
}

// ComputeBlobProof returns the KZG proof that is used to verify the blob against the commitment
// (types/blobs.go:111-117): the opening at the Fiat-Shamir challenge of EIP-4844 compute_challenge.
func (b *Blob) ComputeBlobProof(commitment KZGCommitment) (KZGProof, error) {
	hs := sha256.New()
	hs.Write([]byte("FSBLOBVERIFY_V1_"))
	var deg [16]byte
	deg[14], deg[15] = 0x10, 0x00 // FIELD_ELEMENTS_PER_BLOB = 4096 as a 16-byte big-endian integer
	hs.Write(deg[:])
	hs.Write(b[:])
	hs.Write(commitment[:])
	z := new(big.Int).SetBytes(hs.Sum(nil))
	z.Mod(z, blsModulus)
	proof, _, err := b.ComputeProof(z)
	return proof, err
}
