"""gnark-crypto memory layouts <-> Python integers (host-side marshalling only).

Elements are `[L]uint64` little-endian limbs in Montgomery form x*2^(64L) mod q; G1Affine is
{X, Y}; G2Affine is {X{A0,A1}, Y{A0,A1}} (BW6-761: {X, Y} over Fp); infinity is all-zero
(SURVEY.md A.4).  These helpers are what the Python mirror of the prover interface uses to hand
buffers to the C ABI - the Go shim passes gnark's own slices instead and needs none of this.
"""
import numpy as np

# (p, r, fp limbs64, fr limbs64, g2 extension degree)
CURVES = {
    1: ("bn254",
        0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
        0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001, 4, 4, 2),
    2: ("bls12_377",
        0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
        0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001, 6, 4, 2),
    3: ("bls12_381",
        0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
        0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 6, 4, 2),
    4: ("bw6_761",
        0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b,
        0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001, 12, 6, 1),
}
NAME_TO_ID = {v[0]: k for k, v in CURVES.items()}


class Layout:
    def __init__(self, curve_id):
        if isinstance(curve_id, str):
            curve_id = NAME_TO_ID[curve_id]
        self.id = curve_id
        self.name, self.p, self.r, self.fp_l, self.fr_l, self.g2_deg = CURVES[curve_id]
        self.fp_bytes = 8 * self.fp_l
        self.fr_bytes = 8 * self.fr_l
        self.Rp = (1 << (64 * self.fp_l)) % self.p
        self.Rr = (1 << (64 * self.fr_l)) % self.r
        self.Rp_inv = pow(self.Rp, -1, self.p)
        self.Rr_inv = pow(self.Rr, -1, self.r)

    # ---- scalars / base field
    def enc_fr(self, vals, mont=True):
        out = bytearray()
        for v in vals:
            v = int(v) % self.r
            if mont:
                v = v * self.Rr % self.r
            out += v.to_bytes(self.fr_bytes, "little")
        return np.frombuffer(bytes(out), dtype=np.uint8).copy()

    def dec_fr(self, buf, mont=True):
        b = bytes(np.asarray(buf, dtype=np.uint8))
        n = len(b) // self.fr_bytes
        vals = [int.from_bytes(b[i * self.fr_bytes:(i + 1) * self.fr_bytes], "little") for i in range(n)]
        return [v * self.Rr_inv % self.r for v in vals] if mont else vals

    def enc_fp(self, vals, mont=True):
        out = bytearray()
        for v in vals:
            v = int(v) % self.p
            if mont:
                v = v * self.Rp % self.p
            out += v.to_bytes(self.fp_bytes, "little")
        return np.frombuffer(bytes(out), dtype=np.uint8).copy()

    def dec_fp(self, buf, mont=True):
        b = bytes(np.asarray(buf, dtype=np.uint8))
        n = len(b) // self.fp_bytes
        vals = [int.from_bytes(b[i * self.fp_bytes:(i + 1) * self.fp_bytes], "little") for i in range(n)]
        return [v * self.Rp_inv % self.p for v in vals] if mont else vals

    # ---- coordinates of a group (1 = G1, 2 = G2): flat list of base-field ints per coordinate
    def coord_width(self, group):
        return 1 if (group == 1 or self.g2_deg == 1) else 2

    def _flat(self, coord, group):
        if self.coord_width(group) == 1:
            return [coord]
        return [coord[0], coord[1]]

    def _unflat(self, vals, group):
        if self.coord_width(group) == 1:
            return vals[0]
        return (vals[0], vals[1])

    def affine_bytes(self, group):
        return 2 * self.coord_width(group) * self.fp_bytes

    def xyzz_bytes(self, group):
        return 4 * self.coord_width(group) * self.fp_bytes

    def enc_affine(self, pts, group):
        flat = []
        w = self.coord_width(group)
        for pt in pts:
            if pt is None:
                flat += [0] * (2 * w)
            else:
                flat += self._flat(pt[0], group) + self._flat(pt[1], group)
        # infinity must stay all-zero (not Montgomery-encoded zero, which is also zero) - fine
        return self.enc_fp(flat)

    def dec_affine(self, buf, group):
        w = self.coord_width(group)
        vals = self.dec_fp(buf)
        pts = []
        for i in range(0, len(vals), 2 * w):
            c = vals[i:i + 2 * w]
            if all(v == 0 for v in c):
                pts.append(None)
            else:
                pts.append((self._unflat(c[:w], group), self._unflat(c[w:], group)))
        return pts

    def enc_xyzz(self, pts, group):
        """pts: list of (X, Y, ZZ, ZZZ) coordinate tuples (infinity: ZZ == 0)."""
        flat = []
        for pt in pts:
            for coord in pt:
                flat += self._flat(coord, group)
        return self.enc_fp(flat)

    def dec_xyzz(self, buf, group):
        w = self.coord_width(group)
        vals = self.dec_fp(buf)
        pts = []
        for i in range(0, len(vals), 4 * w):
            c = vals[i:i + 4 * w]
            pts.append(tuple(self._unflat(c[k * w:(k + 1) * w], group) for k in range(4)))
        return pts
