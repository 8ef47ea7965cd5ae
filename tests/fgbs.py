"""Reader/writer of the tiny "FGBS" state container exchanged with oracle/_ref/ref_sim
(the reference's own CUDA build).  TEST INFRASTRUCTURE."""
import json
import os
import struct
import subprocess

import numpy as np

MAGIC = 0x53424746
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SIM = os.path.join(ROOT, "oracle", "_ref", "ref_sim")
# the reference with four method bodies replaced by integration/fgb_reference_shims.cu (oracle/ref_build/build_dropin.sh)
REF_SIM_FGB = os.path.join(ROOT, "oracle", "_ref", "ref_sim_fgb")


def write_state(path, columns):
    """columns: dict name -> np.ndarray [n] or [n, elements]"""
    n = len(next(iter(columns.values()))) if columns else 0
    with open(path, "wb") as f:
        f.write(struct.pack("<III", MAGIC, len(columns), n))
        for name, a in columns.items():
            a = np.ascontiguousarray(a)
            elements = 1 if a.ndim == 1 else a.shape[1]
            f.write(name.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<II", a.dtype.itemsize, elements))
            f.write(a.tobytes())


def read_state(path, dtypes=None):
    """returns dict name -> array (uint32 view unless dtypes[name] says otherwise)"""
    dtypes = dtypes or {}
    out = {}
    with open(path, "rb") as f:
        magic, nvars, n = struct.unpack("<III", f.read(12))
        assert magic == MAGIC, path
        for _ in range(nvars):
            name = f.read(32).split(b"\0")[0].decode()
            es, el = struct.unpack("<II", f.read(8))
            raw = f.read(n * es * el)
            dt = dtypes.get(name, {4: np.uint32, 8: np.uint64, 1: np.uint8, 2: np.uint16}[es])
            a = np.frombuffer(raw, dtype=dt).copy()
            out[name] = a if el == 1 else a.reshape(n, el)
    return out


def have_ref():
    return os.path.exists(REF_SIM) and os.access(REF_SIM, os.X_OK)


def have_dropin():
    return have_ref() and os.path.exists(REF_SIM_FGB) and os.access(REF_SIM_FGB, os.X_OK)


def run_ref(model, params, in_state, out_prefix, steps=1, warmup=0, dump_messages=None, timeout=600, pops=None, dumps=None,
            dump_steps=False, binary=None):
    """Runs the reference's CUDA build; returns the parsed JSON line it prints.
    in_state: state file of the model's main agent (or None); pops: [(agent, state, path)] further populations;
    dumps: [(agent, state)] -> <out_prefix>.<agent>.<state>.bin; dump_steps: <out_prefix>.s<k>.<agent>.bin after each step."""
    cmd = [binary or REF_SIM, "--model", model, "--params", ",".join(f"{k}={v}" for k, v in params.items()), "--out",
           out_prefix, "--steps", str(steps), "--warmup", str(warmup), "--quiet"]
    if in_state:
        cmd += ["--in", in_state]
    for a, st, path in pops or []:
        cmd += ["--pop", f"{a}:{st}:{path}"]
    for a, st in dumps or []:
        cmd += ["--dump", f"{a}:{st}" if st else a]
    if dump_steps:
        cmd += ["--dump-steps"]
    if dump_messages:
        cmd += ["--dump-messages", dump_messages]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_sim failed ({r.returncode}): {r.stderr[-2000:]}\n{r.stdout[-500:]}")
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("ref_sim printed no JSON: " + r.stdout[-500:])
