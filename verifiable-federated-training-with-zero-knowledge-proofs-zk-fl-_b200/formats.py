"""File formats on the drop-in boundary: iden3 binfile containers (`.wtns`, `.zkey`,
`.r1cs`), and the snarkjs JSON shapes `proof.json`, `public.json`, `verification_key.json`."""
from __future__ import annotations

import struct

FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_RQ_INV = pow(1 << 256, -1, FQ)


def write_container(magic: bytes, version: int, sections) -> bytes:
    out = [magic, struct.pack("<II", version, len(sections))]
    for sid, payload in sections:
        out += [struct.pack("<IQ", sid, len(payload)), payload]
    return b"".join(out)


def read_container(data: bytes, magic: bytes) -> dict:
    if data[:4] != magic:
        raise ValueError(f"not a {magic.decode()} file")
    n = struct.unpack_from("<I", data, 8)[0]
    pos, out = 12, {}
    for _ in range(n):
        sid, ln = struct.unpack_from("<IQ", data, pos)
        pos += 12
        out.setdefault(sid, data[pos:pos + ln])
        pos += ln
    return out


# ---------------------------------------------------------------- .wtns (v2)
def wtns_write(values_le: bytes) -> bytes:
    """values_le: n_witness * 32 bytes, canonical little-endian."""
    n = len(values_le) // 32
    hdr = struct.pack("<I", 32) + FR.to_bytes(32, "little") + struct.pack("<I", n)
    return write_container(b"wtns", 2, [(1, hdr), (2, values_le)])


def wtns_read(data: bytes) -> bytes:
    s = read_container(data, b"wtns")
    n8 = struct.unpack_from("<I", s[1], 0)[0]
    if n8 != 32 or int.from_bytes(s[1][4:36], "little") != FR:
        raise ValueError("wtns: not over the BN254 scalar field")
    n = struct.unpack_from("<I", s[1], 36)[0]
    if len(s[2]) != 32 * n:
        raise ValueError("wtns: truncated")
    return s[2]


# ---------------------------------------------------------------- .zkey header / vkey
def _g1_mont_to_dec(b: bytes):
    if b == bytes(64):
        return ["0", "1", "0"]
    x = int.from_bytes(b[:32], "little") * _RQ_INV % FQ
    y = int.from_bytes(b[32:], "little") * _RQ_INV % FQ
    return [str(x), str(y), "1"]


def _g2_mont_to_dec(b: bytes):
    if b == bytes(128):
        return [["0", "0"], ["1", "0"], ["0", "0"]]
    v = [int.from_bytes(b[32 * i:32 * i + 32], "little") * _RQ_INV % FQ for i in range(4)]
    return [[str(v[0]), str(v[1])], [str(v[2]), str(v[3])], ["1", "0"]]


def zkey_header(data: bytes) -> dict:
    s = read_container(data, b"zkey")
    h = s[2]
    n_vars, n_public, domain = struct.unpack_from("<III", h, 72)
    return {"n_vars": n_vars, "n_public": n_public, "domain": domain, "sections": s}


def export_verification_key(zkey: bytes) -> dict:
    """`snarkjs zkey export verificationkey` (tests/full_system_simulation.mjs:733-735).
    vk_alphabeta_12 is omitted: snarkjs' verifier recomputes nothing from it and neither do we."""
    info = zkey_header(zkey)
    h = info["sections"][2]
    p = 84
    alpha1 = h[p:p + 64]; p += 64
    p += 64
    beta2 = h[p:p + 128]; p += 128
    gamma2 = h[p:p + 128]; p += 128
    p += 64
    delta2 = h[p:p + 128]
    ic = info["sections"][3]
    return {
        "protocol": "groth16", "curve": "bn128", "nPublic": info["n_public"],
        "vk_alpha_1": _g1_mont_to_dec(alpha1), "vk_beta_2": _g2_mont_to_dec(beta2),
        "vk_gamma_2": _g2_mont_to_dec(gamma2), "vk_delta_2": _g2_mont_to_dec(delta2),
        "IC": [_g1_mont_to_dec(ic[64 * i:64 * i + 64]) for i in range(len(ic) // 64)],
    }


# ---------------------------------------------------------------- proof.json / public.json
def proof_bytes_to_json(p: bytes) -> dict:
    v = [str(int.from_bytes(p[32 * i:32 * i + 32], "little")) for i in range(8)]
    return {"pi_a": [v[0], v[1], "1"], "pi_b": [[v[2], v[3]], [v[4], v[5]], ["1", "0"]],
            "pi_c": [v[6], v[7], "1"], "protocol": "groth16", "curve": "bn128"}


def proof_json_to_bytes(j: dict) -> bytes:
    """proof.json -> 256 bytes.  snarkjs reads the points as projective triples with z = 1 (`["1","0"]` in G2) and only proves
    / verifies `groth16` over `bn128`: anything else is malformed here (ValueError), not silently reinterpreted."""
    if j.get("protocol", "groth16") != "groth16" or j.get("curve", "bn128") not in ("bn128", "bn254"):
        raise ValueError("proof.json: not a groth16 proof over bn128")
    for key in ("pi_a", "pi_c"):
        if len(j[key]) != 3 or str(j[key][2]) != "1":
            raise ValueError(f"proof.json: {key} is not an affine point (z != 1)")
    if len(j["pi_b"]) != 3 or [str(v) for v in j["pi_b"][2]] != ["1", "0"]:
        raise ValueError("proof.json: pi_b is not an affine point (z != 1)")
    vals = [j["pi_a"][0], j["pi_a"][1], j["pi_b"][0][0], j["pi_b"][0][1], j["pi_b"][1][0], j["pi_b"][1][1],
            j["pi_c"][0], j["pi_c"][1]]
    return b"".join(int(x).to_bytes(32, "little") for x in vals)


def publics_bytes_to_json(p: bytes) -> list:
    return [str(int.from_bytes(p[i:i + 32], "little")) for i in range(0, len(p), 32)]


# ---------------------------------------------------------------- verification key <-> bytes
def _g1_dec_to_bytes(p) -> bytes:
    if p[2] == "0":
        return bytes(64)
    return int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little")


def _g2_dec_to_bytes(p) -> bytes:
    if p[2][0] == "0" and p[2][1] == "0":
        return bytes(128)
    return b"".join(int(v).to_bytes(32, "little") for v in (p[0][0], p[0][1], p[1][0], p[1][1]))


def vkey_json_to_bytes(vk: dict) -> dict:
    """verification_key.json -> the byte fields zkfl_groth16_verify takes (affine canonical little-endian)."""
    return {"alpha1": _g1_dec_to_bytes(vk["vk_alpha_1"]), "beta2": _g2_dec_to_bytes(vk["vk_beta_2"]),
            "gamma2": _g2_dec_to_bytes(vk["vk_gamma_2"]), "delta2": _g2_dec_to_bytes(vk["vk_delta_2"]),
            "ic": b"".join(_g1_dec_to_bytes(p) for p in vk["IC"]), "n_public": int(vk["nPublic"])}
