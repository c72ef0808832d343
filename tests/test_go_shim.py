"""The Go shim cannot be compiled here (no Go toolchain), so it is checked structurally: every `C.b200_*` call in
davinci-node_b200/go names a function declared in include/b200_groth16.h and passes as many arguments as the
prototype has; every field set in a C struct literal exists in the header's struct; the generated per-curve files are
in sync with their template; the exported Go surface matches the reference's prover package."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "davinci-node_b200", "go")
HEADER = open(os.path.join(ROOT, "include", "b200_groth16.h")).read()


def header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    protos = {}
    for m in re.finditer(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_args(args))
    return protos


def header_struct_fields():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(b200_\w+);", text, flags=re.S):
        fields = set()
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                name = re.findall(r"(\w+)\s*$", part.strip())
                if name:
                    fields.add(name[0])
        out[m.group(2)] = fields
    return out


def _split_args(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def go_files():
    for d, _, files in os.walk(GO):
        for f in files:
            if f.endswith(".go"):
                yield os.path.join(d, f)


def _call_args(text, start):
    """argument string of the call whose '(' is at text[start]"""
    depth, i = 0, start
    while True:
        ch = text[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return text[start + 1:i]
        i += 1


def test_every_c_call_matches_the_header():
    protos = header_prototypes()
    assert len(protos) >= 40
    seen = set()
    for path in go_files():
        text = open(path).read()
        text = re.sub(r"//[^\n]*", "", text)
        for m in re.finditer(r"C\.(b200_\w+)\s*\(", text):
            name = m.group(1)
            assert name in protos, "%s calls %s, which include/b200_groth16.h does not declare" % (path, name)
            args = _call_args(text, m.end() - 1).strip()
            n = 0 if not args else len(_split_args(args))
            assert n == protos[name], "%s: %s called with %d arguments, header has %d" % (path, name, n, protos[name])
            seen.add(name)
    for must in ("b200_init", "b200_pk_register", "b200_commit", "b200_prove", "b200_pk_release", "b200_kzg_srs_register",
                 "b200_kzg_srs_add_monomial", "b200_blob_commit", "b200_blob_cell_proofs", "b200_blob_proof", "b200_last_error"):
        assert must in seen, must


def test_struct_literals_use_header_fields():
    fields = header_struct_fields()
    for path in go_files():
        text = open(path).read()
        for m in re.finditer(r"C\.(b200_pk_desc|b200_prove_in|b200_proof_out)\s*\{", text):
            body = _call_args(text.replace("{", "(").replace("}", ")"), m.end() - 1)
            names = re.findall(r"(?:^|[\s,])(\w+):", body)
            assert names, path
            for n in names:
                assert n in fields[m.group(1)], "%s: field %s is not in %s" % (path, n, m.group(1))
    # every field of the prove input is set (a forgotten field would silently be zero)
    tmpl = open(os.path.join(GO, "prover", "curve.go.tmpl")).read()
    body = tmpl[tmpl.index("in := C.b200_prove_in{"):tmpl.index("for _, sl := range []C.b200_slice{in.wires")]
    for f in fields["b200_prove_in"]:
        assert re.search(r"\b%s:" % f, body), f


def test_generated_curve_files_are_in_sync():
    sys.path.insert(0, GO)
    import gen_curves
    outs = gen_curves.outputs()
    assert len(outs) == 4
    for path, text in outs.items():
        assert open(path).read() == text, "%s is stale: run davinci-node_b200/go/gen_curves.py" % path
        for enum in re.findall(r"C\.(B200_\w+)", text):
            assert re.search(r"\b%s\s*=" % enum, HEADER), enum


def test_exported_surface_matches_the_reference_package():
    """prover_cpu.go:19-64 / prover_gpu.go:66-164 export these names with these parameter lists."""
    text = open(os.path.join(GO, "prover", "prover_b200.go")).read()
    sig_c = r"\(curveID ecc\.ID, ccs constraint\.ConstraintSystem, pk groth16\.ProvingKey, assignment frontend\.Circuit, opts \.\.\.backend\.ProverOption\) \(groth16\.Proof, error\)"
    sig_w = r"\(curveID ecc\.ID, ccs constraint\.ConstraintSystem, pk groth16\.ProvingKey, w witness\.Witness, opts \.\.\.backend\.ProverOption\) \(groth16\.Proof, error\)"
    for name in ("Prove", "CPUProver", "GPUProver"):
        assert re.search(r"func %s%s" % (name, sig_c), text), name
    for name in ("ProveWithWitness", "CPUProverWithWitness", "GPUProverWithWitness"):
        assert re.search(r"func %s%s" % (name, sig_w), text), name
    for curve in ("BN254", "BLS12_377", "BLS12_381", "BW6_761"):
        assert "case ecc.%s:" % curve in text
        assert "proving key type mismatch for %s" % curve in text
    assert text.count("{") == text.count("}") and text.count("(") == text.count(")")
    blobs = open(os.path.join(GO, "types", "blobs_b200.go")).read()
    for sig in ("func (b *Blob) ComputeCommitment() (KZGCommitment, error)",
                "func (b *Blob) ComputeCellProofs() ([]KZGProof, error)",
                "func (b *Blob) ComputeBlobProof(commitment KZGCommitment) (KZGProof, error)",
                "func (b *Blob) ComputeProof(point *big.Int) (proof KZGProof, claim *big.Int, err error)"):
        assert sig in blobs, sig
    assert blobs.count("{") == blobs.count("}") and blobs.count("(") == blobs.count(")")


def _strip_go(text):
    """Go source with comments, string / rune literals and raw strings blanked out (positions preserved)."""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            j = text.find("\n", i)
            j = n if j < 0 else j
            out.append(" " * (j - i))
            i = j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            j = n if j < 0 else j + 2
            out.append("".join(ch if ch == "\n" else " " for ch in text[i:j]))
            i = j
        elif c in "\"'`":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if (text[j] == "\\" and c != "`") else 1
            out.append(c + " " * (j - i - 1) + c)
            i = j + 1
        else:
            out.append(c)
            i += 1
    return "".join(out)


def test_go_files_are_lexically_balanced_and_declare_what_they_use():
    """No Go toolchain here: at least every bracket closes, every import alias is used, and every function the
    dispatcher calls (prove<ID> / verify<ID>) is defined by a generated file."""
    import glob
    import re
    gofiles = sorted(glob.glob(os.path.join(GO, "prover", "*.go")) + glob.glob(os.path.join(GO, "types", "*.go")))
    assert gofiles
    defined, called = set(), set()
    for path in gofiles:
        src = _strip_go(open(path).read())
        stack = []
        pairs = {")": "(", "]": "[", "}": "{"}
        for pos, ch in enumerate(src):
            if ch in "([{":
                stack.append((ch, pos))
            elif ch in ")]}":
                assert stack and stack[-1][0] == pairs[ch], "%s: unbalanced %r at offset %d" % (os.path.basename(path), ch, pos)
                stack.pop()
        assert not stack, "%s: unclosed %r" % (os.path.basename(path), stack[-1])
        # import aliases / package names must be referenced
        m = re.search(r"\bimport\s*\(\s*(.*?)\)", open(path).read().split('import "C"')[-1], re.S)
        if m:
            for line in m.group(1).splitlines():
                line = line.strip()
                if not line or line.startswith("//"):
                    continue
                parts = line.split()
                pkg = parts[-1].strip('"')
                alias = parts[0] if len(parts) == 2 else pkg.split("/")[-1]
                body = src[src.index(m.group(1)[:20]) + len(m.group(1)):] if m.group(1)[:20] in src else src
                assert re.search(r"\b%s\." % re.escape(alias), body), "%s: import %s (%s) unused" % (os.path.basename(path), alias, pkg)
        defined |= set(re.findall(r"^func\s+(\w+)\s*\(", src, re.M))
        called |= set(re.findall(r"\b((?:prove|verify|register)(?:BN254|BLS12377|BLS12381|BW6761))\s*\(", src))
    assert called and called <= defined, called - defined
