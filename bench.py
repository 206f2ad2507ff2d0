#!/usr/bin/env python3
"""bench.py — throughput of the bee2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--paths belt_ctr,bash512,bign_verify]

One JSON line on stdout (rank 0). The headline (`metric`/`value`) is BASELINE.json configs[1]:
belt-CTR, one key, 1 GiB of keystream per GPU per step. The other two parts of BASELINE.json's
metric (bash-512 on 2^20 x 4 KiB messages, bign-curve256v1 verify on 2^18 signatures) are measured
in the same run: in full under `"paths"`, and compactly (value, e2e, issue / HBM fraction, CPU
baseline) under `config.also` and again as the LAST key `also` of the line. `--impl reference`
times the same three paths and prints them in the same places.

N > 1 additionally measures STRONG scaling of the fixed-size BASELINE configs (2^26 CTR blocks,
2^20 messages, 2^18 signatures, 2^26 (key, block) pairs) sharded N ways, under `config.strong`:
compute only, compute with the final gather fused into the kernel (each rank's kernel stores its
outputs straight into rank 0's HBM over NVLink through a CUDA-IPC mapping), and compute followed by
an NCCL gather; rank 0 checks that the gathered bytes equal its own single-GPU run of the whole batch.

A step = one pass of the path over one batch. `value` is timed with CUDA events on the stream the
kernels are launched on, inputs resident in HBM; `e2e` goes through the host-pointer C-ABI call
(pinned host buffers, H2D and D2H inside the timed region). N > 1: one process per GPU (torchrun),
every rank runs the same per-GPU batch (weak scaling; rank r takes counter blocks / items
[r*units, (r+1)*units)), key / iv / oid are broadcast from rank 0 over NCCL, timing is the max over
ranks. `--impl reference` times the reference's own CPU code (oracle/_ref) on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
OID = bytes.fromhex("06092A7000020022651F51")

CFG = {
    "belt_ctr": dict(metric="belt-CTR keystream GB/s", unit="GB/s", workload="belt-CTR single key, 1 GiB keystream per GPU",
                     units=1 << 26, unit_bytes=16, dtype="u32"),
    "bash512": dict(metric="bash-512 GB/s hashed", unit="GB/s", workload="bash-512 batch: 2^20 messages x 4 KiB per GPU",
                    units=1 << 20, unit_bytes=4096, dtype="u64"),
    "belt_dwp": dict(metric="belt-DWP wrap GB/s", unit="GB/s",
                     workload="belt-DWP authenticated encryption: 1 GiB critical + 4 KiB open data per GPU, one key",
                     units=1 << 26, unit_bytes=16, dtype="u32"),
    "belt_ecb": dict(metric="belt-ECB key-agility GB/s", unit="GB/s",
                     workload="belt-ECB key agility: 2^26 blocks under 2^26 independent 32-byte keys per GPU",
                     units=1 << 26, unit_bytes=16, dtype="u32"),
    "bign_verify": dict(metric="bign-curve256v1 verifies/s", unit="verifies/s",
                        workload="bign-curve256v1 batch verify: 2^18 signatures per GPU (1/16 corrupted)",
                        units=1 << 18, unit_bytes=148, dtype="u32"),
    "bign_sign2": dict(metric="bign-curve256v1 deterministic signatures/s", unit="signatures/s",
                       workload="bign-curve256v1 batch bignSign2: 2^18 (hash, private key) pairs per GPU",
                       units=1 << 18, unit_bytes=112, dtype="u32"),
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the bench
# configuration, from the committed `ncu --set full` captures (profiles/r02_ncu_raw_*.csv)
NCU_BIGN_WIDE_PER_VERIFY = 96884    # IMAD.WIDE(.X) executed per verify (profiles/r02_bign_opcode_mix.json)
NCU_TRAFFIC = {"bign_sign2": 4.2928e+08, "belt_dwp": 1.0774e+09, "belt_ecb": 4.2638e+09, "belt_ctr": 1.0135e+09, "bash512": 4.4449e+09, "bign_verify": 3.1637e+09}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_has(flag):
    try:
        with open("/proc/cpuinfo") as f:
            return flag in f.read()
    except Exception:
        return False


# ---------------------------------------------------------------- CPU arm (reference / port)
class CpuArm:
    """The reference's CPU implementation (oracle/_ref/libbee2ref_*.so, kind "reference") or, if it
    was not built, our C restatement (kind "port"), fanned out over all host threads by
    oracle/cpu_harness.c."""

    def __init__(self):
        self.h = C.CDLL(os.path.join(REF_DIR, "libcpuharness.so"))
        for n in ("harness_bash", "harness_belt_ctr", "harness_belt_dwp", "harness_belt_ecb_multikey", "harness_bign_verify",
                  "harness_bign_sign2", "harness_bign_pubkey"):
            getattr(self.h, n).restype = C.c_double
        self.threads = host_threads()
        ref64 = os.path.join(REF_DIR, "libbee2ref_64.so")
        self.kind = "reference" if os.path.exists(ref64) else "port"
        if self.kind == "reference":
            self.lib = ref64
            self.bash_lib, self.bash_name = ref64, "BASH_64"
            for flag, suffix, name in (("avx512f", "avx512", "BASH_AVX512"), ("avx2", "avx2", "BASH_AVX2")):
                p = os.path.join(REF_DIR, f"libbee2ref_{suffix}.so")
                if cpu_has(flag) and os.path.exists(p):
                    self.bash_lib, self.bash_name = p, name
                    break
        else:
            self.lib = self.bash_lib = os.path.join(REF_DIR, "libbee2oracle.so")
            self.bash_name = "port"
        self.is_port = int(self.kind == "port")

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def bash(self, msgs):
        out = np.zeros((msgs.shape[0], 64), dtype=np.uint8)
        dt = self.h.harness_bash(self.bash_lib.encode(), self.is_port, self._p(out), C.c_size_t(256), self._p(msgs),
                                 C.c_size_t(msgs.shape[1]), C.c_size_t(msgs.shape[1]), C.c_size_t(msgs.shape[0]), self.threads)
        return dt, out

    def belt_ctr(self, buf, key, iv):
        unit = 1 << 16
        dt = self.h.harness_belt_ctr(self.lib.encode(), self.is_port, self._p(buf), None, C.c_size_t(unit),
                                     C.c_size_t(buf.size // unit), key, iv, self.threads)
        return dt

    def belt_dwp(self, buf, key, iv):
        unit = 1 << 16
        return self.h.harness_belt_dwp(self.lib.encode(), self.is_port, self._p(buf), C.c_size_t(unit),
                                       C.c_size_t(buf.size // unit), key, iv, self.threads)

    def ecb_multikey(self, blocks, keys):
        return self.h.harness_belt_ecb_multikey(self.lib.encode(), self.is_port, self._p(blocks), self._p(keys),
                                                C.c_size_t(blocks.shape[0]), self.threads)

    def verify(self, hashes, sigs, pubs):
        st = np.zeros(hashes.shape[0], dtype=np.uint32)
        dt = self.h.harness_bign_verify(self.lib.encode(), self.is_port, self._p(st), self._p(hashes), self._p(sigs),
                                        self._p(pubs), C.c_size_t(hashes.shape[0]), self.threads)
        return dt, st

    def sign2(self, hashes, privs):
        st = np.zeros(hashes.shape[0], dtype=np.uint32)
        sigs = np.zeros((hashes.shape[0], 48), dtype=np.uint8)
        dt = self.h.harness_bign_sign2(self.lib.encode(), self.is_port, self._p(st), self._p(sigs), self._p(hashes),
                                       self._p(privs), C.c_size_t(hashes.shape[0]), self.threads)
        return dt, st, sigs

    def pubkey(self, privs):
        st = np.zeros(privs.shape[0], dtype=np.uint32)
        pubs = np.zeros((privs.shape[0], 64), dtype=np.uint8)
        dt = self.h.harness_bign_pubkey(self.lib.encode(), self.is_port, self._p(st), self._p(pubs), self._p(privs),
                                        C.c_size_t(privs.shape[0]), self.threads)
        return dt, st, pubs

    # one bounded sample of a path: returns (seconds, units processed, description)
    def sample(self, path, target_s, state):
        rng = state.setdefault("rng", np.random.default_rng(1))
        rate = state.get(("rate", path))
        if path == "belt_ctr":
            nbytes = (4 << 20) * self.threads if rate is None else int(min(max(rate * target_s, 1 << 22), 1 << 30))
            nbytes -= nbytes % (1 << 16)
            buf = state.get(("ctrbuf", nbytes))
            if buf is None:
                buf = state[("ctrbuf", nbytes)] = np.zeros(nbytes, dtype=np.uint8)
            dt = self.belt_ctr(buf, bytes(range(32)), bytes(16))
            state["last"] = buf
            units, desc = nbytes, f"{nbytes >> 20} MiB keystream, {self.threads} independent beltCTR shards"
        elif path == "belt_dwp":
            nbytes = (4 << 20) * self.threads if rate is None else int(min(max(rate * target_s, 1 << 22), 1 << 30))
            nbytes -= nbytes % (1 << 16)
            buf = state.get(("dwpbuf", nbytes))
            if buf is None:
                buf = state[("dwpbuf", nbytes)] = np.zeros(nbytes, dtype=np.uint8)
            dt = self.belt_dwp(buf, bytes(range(32)), bytes(16))
            state["last"] = buf
            units, desc = nbytes, f"{nbytes >> 20} MiB, {self.threads} independent beltDWPWrap shards"
        elif path == "belt_ecb":
            n = (1 << 14) * self.threads if rate is None else int(min(max(rate * target_s / 16, 1 << 14), 1 << 26))
            pair = state.get(("ecb", n))
            if pair is None:
                pair = state[("ecb", n)] = (rng.integers(0, 256, (n, 16), dtype=np.uint8),
                                            rng.integers(0, 256, (n, 32), dtype=np.uint8))
            dt = self.ecb_multikey(*pair)
            state["last"] = pair
            units, desc = n * 16, f"{n} (key, block) pairs via beltECBEncr(16 B, key_i)"
        elif path == "bash512":
            n = 256 * self.threads if rate is None else int(min(max(rate * target_s / 4096, 64), 1 << 20))
            msgs = state.get(("bashmsgs", n))
            if msgs is None:
                msgs = state[("bashmsgs", n)] = rng.integers(0, 256, (n, 4096), dtype=np.uint8)
            dt, _ = self.bash(msgs)
            state["last"] = msgs
            units, desc = n * 4096, f"{n} messages x 4 KiB via bashHash ({self.bash_name})"
        elif path == "bign_sign2":
            n = 16 * self.threads if rate is None else int(min(max(rate * target_s, 16), 1 << 18))
            key = ("keys", n)
            if key not in state:
                priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
                priv[:, 31] &= 0x7F
                state[key] = (rng.integers(0, 256, (n, 32), dtype=np.uint8), priv)
            dt, st, _ = self.sign2(*state[key])
            state["last"] = state[key]
            assert not st.any()
            units, desc = n, f"{n} signatures via " + ("bign128Sign2" if not self.is_port else "orc_bignSign2_128")
        else:
            n = 16 * self.threads if rate is None else int(min(max(rate * target_s, 16), 1 << 18))
            key = ("sigs", n)
            if key not in state:
                priv = rng.integers(0, 256, (n, 32), dtype=np.uint8)
                priv[:, 31] &= 0x7F
                hashes = rng.integers(0, 256, (n, 32), dtype=np.uint8)
                _, st, pubs = self.pubkey(priv)
                _, st2, sigs = self.sign2(hashes, priv)
                assert not st.any() and not st2.any()
                bad = np.arange(0, n, 16)            # the GPU leg's corruption pattern
                sigs[bad, bad % 48] ^= 1
                state[key] = (hashes, sigs, pubs)
            hashes, sigs, pubs = state[key]
            dt, st = self.verify(hashes, sigs, pubs)
            state["last"] = (hashes, sigs, pubs)
            bad = np.arange(0, n, 16)
            assert (st[bad] == 510).all() and int((st == 0).sum()) == n - len(bad), "cpu verify statuses off"
            units, desc = n, f"{n} signatures (1/16 corrupted) via " + ("bign128Verify" if not self.is_port else "orc_bignVerify128")
        if dt <= 0:
            raise RuntimeError(f"cpu harness failed for {path}: {dt}")
        # the batch is capped at the config size: repeat it until the sample lasts about target_s
        reps = 1
        while rate is not None and dt * reps < 0.75 * target_s and reps < 64:
            again = self.sample_once(path, state)
            if again <= 0:
                raise RuntimeError(f"cpu harness failed for {path}: {again}")
            dt += again
            reps += 1
        if reps > 1:
            dt, units, desc = dt, units * reps, desc + f" x {reps} passes"
        state[("rate", path)] = units / dt
        return dt, units, desc

    def sample_once(self, path, state):
        last = state["last"]
        if path == "belt_ctr":
            return self.belt_ctr(last, bytes(range(32)), bytes(16))
        if path == "bash512":
            return self.bash(last)[0]
        if path == "belt_ecb":
            return self.ecb_multikey(*last)
        if path == "belt_dwp":
            return self.belt_dwp(last, bytes(range(32)), bytes(16))
        if path == "bign_sign2":
            return self.sign2(*last)[0]
        return self.verify(*last)[0]

    def baseline(self, path, target_s=4.0):
        state = {}
        self.sample(path, 0, state)                 # probe (also warms caches / lazy curve)
        dt, units, desc = self.sample(path, target_s, state)
        scale = 1e9 if CFG[path]["unit"] == "GB/s" else 1.0
        out = {"value": units / dt / scale, "unit": CFG[path]["unit"], "cores": self.threads, "kind": self.kind,
               "sample": desc + f", {dt:.2f} s wall"}
        # the single-thread figure (and the scalar BASH_64 build for bash) alongside
        saved = self.threads, self.bash_lib, self.bash_name
        try:
            self.threads = 1
            st1 = {}
            self.sample(path, 0, st1)
            dt1, u1, _ = self.sample(path, 1.0, st1)
            out["value_1_thread"] = u1 / dt1 / scale
            if path == "bash512" and self.kind == "reference" and saved[2] != "BASH_64":
                self.threads, self.bash_lib, self.bash_name = saved[0], self.lib, "BASH_64"
                st64 = {}
                self.sample(path, 0, st64)
                dt64, u64_, _ = self.sample(path, 1.5, st64)
                out["value_bash_64"] = u64_ / dt64 / scale
        finally:
            self.threads, self.bash_lib, self.bash_name = saved
        return out


def compact(r):
    """One path's result cut down to what the driver-visible summary needs."""
    out = {"value": r.get("value"), "unit": r.get("unit")}
    if r.get("ms_per_step") is not None:
        out["ms"] = round(r["ms_per_step"], 4)
    if r.get("e2e"):
        out["e2e"] = r["e2e"]["value"]
    if r.get("issue_roofline"):
        out["issue_frac"] = r["issue_roofline"].get("frac")
        out["issue_bound"] = r["issue_roofline"].get("bound", "").split(" ")[0]
    if r.get("roofline"):
        out["hbm_frac"] = r["roofline"].get("frac")
    if r.get("cpu_baseline"):
        out["cpu"] = r["cpu_baseline"]["value"]
        out["cpu_cores"] = r["cpu_baseline"]["cores"]
    for k in ("impl", "sample"):
        if r.get(k):
            out[k] = r[k]
    return {k: (round(v, 6) if isinstance(v, float) else v) for k, v in out.items()}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU code (oracle/_ref: bashHash of the BASH_AVX512 build,
    beltCTR, bign128Verify) on all host threads, same metric/config keys. Every path of --paths is
    timed for `steps` bounded samples; the first one is the headline of the line."""
    if rank != 0:
        return
    arm = CpuArm()
    paths = args.paths
    budget = 150.0 / max(1, len(paths) * (args.steps + args.warmup))
    res = {}
    for path in paths:
        cfg = CFG[path]
        scale = 1e9 if cfg["unit"] == "GB/s" else 1.0
        state = {}
        arm.sample(path, 0, state)
        per_step_s = max(0.25, min(3.0, budget))
        for _ in range(args.warmup):
            arm.sample(path, per_step_s, state)
        tot_t = tot_u = 0.0
        desc = ""
        for _ in range(args.steps):
            dt, units, desc = arm.sample(path, per_step_s, state)
            tot_t += dt
            tot_u += units
        res[path] = {"metric": cfg["metric"], "value": tot_u / tot_t / scale, "unit": cfg["unit"],
                     "ms_per_step": 1e3 * tot_t / args.steps, "sample": desc, "impl": "reference"}
    head = paths[0]
    cfg, h = CFG[head], res[head]
    also = {k: compact(v) for k, v in res.items() if k != head}
    line = {"impl": "reference", "metric": cfg["metric"], "value": h["value"], "unit": cfg["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": h["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["workload"], "reference_step": h["sample"], "also": also},
            "cpu_baseline": {"value": h["value"], "unit": cfg["unit"], "cores": arm.threads, "kind": arm.kind,
                             "sample": f"{args.steps} steps: {h['sample']}"},
            "e2e": {"value": h["value"], "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "paths": {k: v for k, v in res.items() if k != head},
            "also": {k: compact(v) for k, v in res.items()}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.active = threading.Event()
        self.stop = threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------- host placement
def numa_bind(index):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated, so the
    staging memory of the e2e legs is node-local (first touch). Returns what was found / done."""
    info = {"gpu_numa_node": None, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                       # 00000000:1b:00.0 -> 0000:1b:00.0
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["host_numa_nodes"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b_ = part.partition("-")
                    cpus.update(range(int(a), int(b_ or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"], info["cpus"] = True, len(cpus)
    except Exception as e:          # no sysfs / nvml in this container: nothing to bind
        info["note"] = type(e).__name__
    return info


class DevBuf:
    """A raw device pointer as a torch uint8 tensor (``torch.as_tensor(DevBuf(p, n), device=...)``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


# ---------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths", default=None,
                    help="comma list; the first one is the headline metric of the JSON line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section (N > 1)")
    args = ap.parse_args()
    if args.paths is None:
        args.paths = ("belt_ctr,bash512,bign_verify" if args.impl == "reference"
                      else "belt_ctr,bash512,bign_verify,belt_ecb,belt_dwp,bign_sign2")
    args.paths = [p for p in args.paths.split(",") if p]
    for p in args.paths:
        if p not in CFG:
            ap.error(f"unknown path {p}")
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import bee2_b200 as b
    from bee2_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    code = b.b2g_init(local_rank)
    if code:
        raise SystemExit(f"b2g_init({local_rank}) -> {code}: {b.b2g_last_error()}")
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = hbm_peak()
    sampler = ClockSampler(local_rank)
    sampler.start()
    numa = numa_bind(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps, warmup, flush):
        """K steps, CUDA events around each launch on the launching stream; returns total seconds
        (max over ranks) and the number of our kernel launches inside the timed region."""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = b.b2g_launch_count()
        sampler.active.set()
        for s, e in ev:
            if flush:
                flush_buf.fill_(1)          # evict the (smaller-than-L2) inputs between steps
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        sampler.active.clear()
        launches = b.b2g_launch_count() - l0
        barrier()
        total = shard.max_over_ranks(sum(s.elapsed_time(e) for s, e in ev) * 1e-3, device=dev)
        return total, launches

    def timed_host(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        barrier()
        return shard.max_over_ranks(total, device=dev)

    def gather_info(local):
        """The final gather of SURVEY §8e: every rank's output to all ranks with one NCCL all-gather
        over NVLink/NVSwitch, timed on the device (max over ranks). Outside the throughput metric."""
        if world == 1:
            return None
        flat = local.reshape(-1).view(torch.uint8)
        dst = torch.empty(world * flat.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(dst, flat)            # warm-up (communicator, buffers)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(dst, flat)
        e1.record()
        torch.cuda.synchronize()
        sec = shard.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device=dev)
        ok = bool(torch.equal(dst[rank * flat.numel():(rank + 1) * flat.numel()], flat))
        del dst
        return {"collective": "ncclAllGather", "bytes_per_rank": int(flat.numel()), "ms": sec * 1e3,
                "GB/s_per_rank_received": (world - 1) * flat.numel() / sec / 1e9, "own_slice_intact": ok}

    def copy_roof():
        """The e2e roof of this box: every rank moves 1 GiB between pinned host memory and its GPU at the
        same time with plain cudaMemcpyAsync (D2H, then H2D); aggregate GB/s over the max rank time."""
        n = 1 << 30
        hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
        dbuf = torch.empty(n, dtype=torch.uint8, device=dev)
        L = b.lib()
        out = {}
        for name, to_dev in (("d2h", 0), ("h2d", 1)):
            dst, src = (dbuf.data_ptr(), hbuf.data_ptr()) if to_dev else (hbuf.data_ptr(), dbuf.data_ptr())
            best = None
            for it in range(4):
                barrier()
                t0 = time.perf_counter()
                assert L.b2g_memcpy_async(dst, src, n, to_dev, stream) == 0
                torch.cuda.synchronize()
                dt = shard.max_over_ranks(time.perf_counter() - t0, device=dev)
                if it and (best is None or dt < best):
                    best = dt
            out[name] = world * n / best / 1e9
        del hbuf, dbuf
        return out

    results = {}
    issue = {}
    roof = None if args.no_e2e else copy_roof()
    if rank == 0:
        for name, kind in (("lop3", 0), ("shf", 1), ("prmt", 2), ("imad", 4), ("imad_wide", 5), ("lds32", 6), ("lop3+imad_wide", 7),
                           ("lop3+imad", 8), ("lop3+ffma", 9), ("imad_hi", 10), ("lop3+lds32", 11), ("ffma", 12), ("dfma", 13), ("dfma+imad_wide", 14),
                           ("imad+imad_wide", 15), ("2iadd.x+imad_wide", 16), ("imad_wide.cc chain", 17),
                           ("iadd3.cc chain", 18), ("imad_wide x8", 19)):
            issue[name] = b.b2g_microbench(kind, 0) / 1e12
    arm = None
    if rank == 0 and not args.no_cpu_baseline:
        arm = CpuArm()

    for path in args.paths:
        cfg = CFG[path]
        units = cfg["units"]
        scale = 1e9 if cfg["unit"] == "GB/s" else 1.0
        e2e_steps = max(1, min(args.steps, 3))
        r = {"metric": cfg["metric"], "unit": cfg["unit"], "workload": cfg["workload"]}
        if path == "belt_ctr":
            secret = np.random.default_rng(10).integers(0, 256, 48, dtype=np.uint8).tobytes() if rank == 0 else None
            kiv = shard.broadcast_bytes(secret, 48, device=dev)        # one NCCL broadcast of key || iv
            st = b.BeltCTR(kiv[:32], kiv[32:])
            key, ctr = st.key_words, st.ctr_words
            nbytes = units * 16
            out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            first = rank * units                      # rank r produces counter blocks [r*units, (r+1)*units)
            fn = lambda: b.beltCTR_dev(out.data_ptr(), 0, nbytes, key, ctr, first, stream)  # noqa: E731
            total, launches = timed(fn, args.steps, args.warmup, flush=False)
            r["l2"] = "output 1 GiB per step > 126 MB L2, no flush needed"
            algo_bytes = nbytes                        # 16 B written per block
            # parity on the spot: head of the stream against the drop-in host call of the same library
            # is checked in tests; here a checksum makes sure the kernel really wrote the buffer
            r["checksum"] = int(out[:: 1 << 16].to(torch.int64).sum().item())
            if not args.no_e2e:
                host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                harr = host.numpy()
                e2e_fn = lambda: b.beltCTRKeystream(nbytes, kiv[:32], kiv[32:], out=harr)  # noqa: E731
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * nbytes * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": 48 + 16, "d2h_bytes_per_step": nbytes + 16,
                            "call": "beltCTRKeystream(dest=pinned host, 1 GiB)", "steps": e2e_steps}
                assert harr[: 1 << 20].tobytes() == out[: 1 << 20].cpu().numpy().tobytes() or rank != 0
                del host, harr
            if world <= 4:                             # 1 GiB per rank: keep the gathered copy under 4 GiB
                r["gather"] = gather_info(out)
            del out
        elif path == "belt_dwp":
            secret = np.random.default_rng(11).integers(0, 256, 48, dtype=np.uint8).tobytes() if rank == 0 else None
            kiv = shard.broadcast_bytes(secret, 48, device=dev)
            st = b.BeltCTR(kiv[:32], kiv[32:])
            key, ctr = st.key_words, st.ctr_words
            nbytes = units * 16
            g = torch.Generator(device=dev).manual_seed(4 + rank)
            data = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=dev, generator=g)
            opn = torch.randint(0, 256, (4096,), dtype=torch.uint8, device=dev, generator=g)
            small = torch.zeros(64, dtype=torch.uint8, device=dev)

            def fn():
                b.beltCTR_dev(data.data_ptr(), data.data_ptr(), nbytes, key, ctr, 0, stream)
                b.beltDWPMac_dev(small.data_ptr(), data.data_ptr(), nbytes, opn.data_ptr(), 4096, key, ctr,
                                 small.data_ptr() + 16, stream)
            # parity on the spot: a 1 MiB prefix through the host entry point of the same library
            pre, po = data[: 1 << 20].cpu().numpy().tobytes(), opn.cpu().numpy().tobytes()
            want = b.beltDWPWrap(pre, po, kiv[:32], kiv[32:])
            chk = data[: 1 << 20].clone()
            b.beltCTR_dev(chk.data_ptr(), chk.data_ptr(), 1 << 20, key, ctr, 0, stream)
            b.beltDWPMac_dev(small.data_ptr(), chk.data_ptr(), 1 << 20, opn.data_ptr(), 4096, key, ctr,
                             small.data_ptr() + 16, stream)
            torch.cuda.synchronize()
            assert (chk.cpu().numpy().tobytes(), small[:8].cpu().numpy().tobytes()) == want
            total, launches = timed(fn, args.steps, args.warmup, flush=False)
            r["l2"] = "data 1 GiB per step > 126 MB L2, no flush needed"
            algo_bytes = units * 48                    # CTR: 16 read + 16 written, tag: 16 read per block
            r["checksum"] = int(small[:8].to(torch.int64).sum().item())
            if not args.no_e2e:
                hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                hin.copy_(data)
                hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                ni, no_ = hin.numpy(), hout.numpy()
                mac = np.zeros(8, dtype=np.uint8)
                L = b.lib()
                kb, ib = np.frombuffer(kiv[:32], dtype=np.uint8).copy(), np.frombuffer(kiv[32:], dtype=np.uint8).copy()
                pon = np.frombuffer(po, dtype=np.uint8).copy()

                def e2e_fn():
                    c = L.beltDWPWrap(no_.ctypes.data, mac.ctypes.data, ni.ctypes.data, nbytes, pon.ctypes.data, 4096,
                                      kb.ctypes.data, 32, ib.ctypes.data)
                    assert c == 0, c
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * nbytes * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": nbytes + 4096, "d2h_bytes_per_step": nbytes + 8,
                            "call": "beltDWPWrap(pinned host 1 GiB -> pinned host 1 GiB + mac)", "steps": e2e_steps}
                del hin, hout, ni, no_
            del data
        elif path == "belt_ecb":
            g = torch.Generator(device=dev).manual_seed(3 + rank)
            keys = torch.randint(0, 256, (units, 32), dtype=torch.uint8, device=dev, generator=g)
            blocks = torch.randint(0, 256, (units, 16), dtype=torch.uint8, device=dev, generator=g)
            first_blocks, first_keys = blocks[:4096].cpu().numpy().copy(), keys[:4096].cpu().numpy().copy()
            fn = lambda: b.beltECBEncrBatch_dev(blocks.data_ptr(), keys.data_ptr(), units, stream)  # noqa: E731
            fn()
            torch.cuda.synchronize()
            # spot parity against the host-pointer entry point of the same library (itself parity-tested)
            assert np.array_equal(blocks[:4096].cpu().numpy(), b.beltECBEncrBatch(first_blocks, first_keys))
            total, launches = timed(fn, args.steps, args.warmup, flush=False)
            r["l2"] = "keys + blocks 3 GiB per step > 126 MB L2, no flush needed"
            algo_bytes = units * 64                    # 32 key + 16 in + 16 out per block
            r["checksum"] = int(blocks[:: 1 << 12].to(torch.int64).sum().item())
            if not args.no_e2e:
                hk = torch.empty((units, 32), dtype=torch.uint8).pin_memory()
                hb = torch.empty((units, 16), dtype=torch.uint8).pin_memory()
                hk.copy_(keys), hb.copy_(blocks)
                nk, nb = hk.numpy(), hb.numpy()
                L = b.lib()

                def e2e_fn():
                    c = L.beltECBEncrBatch(nb.ctypes.data, nk.ctypes.data, units)
                    assert c == 0, c
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * units * 16 * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": units * 48, "d2h_bytes_per_step": units * 16,
                            "call": "beltECBEncrBatch(pinned host blocks 1 GiB + keys 2 GiB, in place)", "steps": e2e_steps}
                del hk, hb, nk, nb
            del keys, blocks
        elif path == "bash512":
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            msgs = torch.randint(0, 256, (units, 4096), dtype=torch.uint8, device=dev, generator=g)
            out = torch.empty((units, 64), dtype=torch.uint8, device=dev)
            fn = lambda: b.bashHashBatch_dev(out.data_ptr(), 256, msgs.data_ptr(), 4096, 4096, units, stream)  # noqa: E731
            total, launches = timed(fn, args.steps, args.warmup, flush=False)
            r["l2"] = "input 4 GiB per step > 126 MB L2, no flush needed"
            algo_bytes = units * 4160                  # 4096 read + 64 written per message
            r["checksum"] = int(out.to(torch.int64).sum().item())
            if not args.no_e2e:
                hmsgs = torch.empty((units, 4096), dtype=torch.uint8).pin_memory()
                hmsgs.copy_(msgs)
                hout = torch.empty((units, 64), dtype=torch.uint8).pin_memory()
                hm, ho = hmsgs.numpy(), hout.numpy()
                L = b.lib()

                def e2e_fn():
                    c = L.bashHashBatch(ho.ctypes.data, 256, hm.ctypes.data, 4096, 4096, units)
                    assert c == 0, c
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * units * 4096 * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": units * 4096, "d2h_bytes_per_step": units * 64,
                            "call": "bashHashBatch(pinned host msgs 4 GiB -> pinned host digests)", "steps": e2e_steps}
                assert np.array_equal(ho[:4096], out[:4096].cpu().numpy())
                del hmsgs, hout, hm, ho
            r["gather"] = gather_info(out)
            del msgs, out
        elif path == "bign_sign2":
            rng = np.random.default_rng(20 + rank)
            priv = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            priv[:, 31] &= 0x7F
            hashes = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            params = b.bignParamsStd()
            oid = shard.broadcast_bytes(OID if rank == 0 else None, len(OID), device=dev)
            d_h, d_k = (torch.from_numpy(x).to(dev) for x in (hashes, priv))
            d_sig = torch.zeros(units * 48, dtype=torch.uint8, device=dev)
            d_st = torch.empty(units, dtype=torch.int32, device=dev)
            fn = lambda: b.bignSign2Batch_dev(d_st.data_ptr(), d_sig.data_ptr(), oid, d_h.data_ptr(), d_k.data_ptr(), units, stream)  # noqa: E731
            total, launches = timed(fn, args.steps, args.warmup, flush=True)
            r["l2"] = "inputs 16 MiB < L2: 256 MiB flush write between steps (outside the events)"
            algo_bytes = units * 112
            assert not d_st.cpu().numpy().any(), "sign2 statuses off"
            # parity on the spot: the signatures verify on the device
            d_pub = torch.zeros(units * 64, dtype=torch.uint8, device=dev)
            b.bignPubkeyCalcBatch_dev(d_st.data_ptr(), d_pub.data_ptr(), d_k.data_ptr(), units, stream)
            b.bignVerifyBatch_dev(d_st.data_ptr(), oid, d_h.data_ptr(), d_sig.data_ptr(), d_pub.data_ptr(), units, stream)
            assert not d_st.cpu().numpy().any(), "signatures do not verify"
            r["checksum"] = int(d_sig[:: 97].to(torch.int64).sum().item())
            if not args.no_e2e:
                ph, pk = (torch.from_numpy(x).pin_memory() for x in (hashes, priv))
                nh, nk = ph.numpy(), pk.numpy()
                pst = torch.empty(units, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
                psig = torch.empty((units, 48), dtype=torch.uint8).pin_memory().numpy()
                e2e_fn = lambda: b.bignSign2Batch(params, oid, nh, nk, status=pst, sigs=psig)  # noqa: E731
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * units * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": units * 64, "d2h_bytes_per_step": units * 52,
                            "call": "bignSign2Batch(pinned host hashes/privkeys -> sigs[], status[])", "steps": e2e_steps}
        else:
            # inputs made by the engine's own batch signer (parity-tested against the oracle), then
            # 1/16 of the items corrupted
            rng = np.random.default_rng(2 + rank)
            priv = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            priv[:, 31] &= 0x7F
            hashes = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            params = b.bignParamsStd()
            st1, pubs = b.bignPubkeyCalcBatch(params, priv)
            st2, sigs = b.bignSign2Batch(params, OID, hashes, priv)
            assert not st1.any() and not st2.any()
            bad = np.arange(0, units, 16)
            sigs[bad, bad % 48] ^= 1
            oid = shard.broadcast_bytes(OID if rank == 0 else None, len(OID), device=dev)
            d_h, d_s, d_p = (torch.from_numpy(x).to(dev) for x in (hashes, sigs, pubs))
            d_st = torch.empty(units, dtype=torch.int32, device=dev)
            fn = lambda: b.bignVerifyBatch_dev(d_st.data_ptr(), oid, d_h.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), units, stream)  # noqa: E731
            total, launches = timed(fn, args.steps, args.warmup, flush=True)
            r["l2"] = "inputs 36 MiB < L2: 256 MiB flush write between steps (outside the events)"
            algo_bytes = units * 148
            stc = d_st.cpu().numpy()
            assert (stc[bad] == 510).all() and int((stc == 0).sum()) == units - len(bad), "verify statuses off"
            r["checksum"] = int(stc.astype(np.int64).sum())
            r["gather"] = gather_info(d_st)
            if not args.no_e2e:
                ph, ps, pp = (torch.from_numpy(x).pin_memory() for x in (hashes, sigs, pubs))
                nh, ns, npb = ph.numpy(), ps.numpy(), pp.numpy()
                pst = torch.empty(units, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
                e2e_fn = lambda: b.bignVerifyBatch(params, oid, nh, ns, npb, status=pst)  # noqa: E731
                t = timed_host(e2e_fn, e2e_steps, 1)
                r["e2e"] = {"value": world * units * e2e_steps / t / scale, "unit": cfg["unit"],
                            "h2d_bytes_per_step": units * 144, "d2h_bytes_per_step": units * 4,
                            "call": "bignVerifyBatch(pinned host hashes/sigs/pubkeys -> status[])", "steps": e2e_steps}
        ms = 1e3 * total / args.steps
        r["value"] = world * units * (cfg["unit_bytes"] if cfg["unit"] == "GB/s" else 1) * args.steps / total / scale
        r["ms_per_step"] = ms
        r["gpu_launches"] = int(launches)
        ach = algo_bytes / (total / args.steps) / 1e9
        r["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": NCU_TRAFFIC.get(path), "traffic_source": "profiles/r02_ncu_raw_*.csv (bytes per launch; belt_dwp: the tag kernel only)",
                         "algorithmic_bytes": algo_bytes, "peak_source": peak_src,
                         "note": "integer-issue bound, not HBM bound (SURVEY §8d): see issue_roofline"}
        results[path] = r


    # ---------------------------------------------------------------- strong scaling (N > 1)
    def strong_section():
        """BASELINE configs 2-5 at their FIXED size, sharded `world` ways by contiguous unit ranges.
        Three timings per path (CUDA events around each rank's work, max over ranks):
          compute      - every rank's kernel on its slice, outputs stay local;
          fused_gather - the same kernel stores its outputs straight into rank 0's HBM over NVLink
                         (CUDA-IPC mapping of rank 0's buffer): compute and the final gather are ONE kernel;
          nccl_gather  - the kernel, then one ncclAllGather of the outputs.
        Rank 0 then runs the WHOLE batch alone and compares the gathered bytes with it."""
        L = b.lib()
        steps, warm = max(3, min(args.steps, 10)), 3
        out = {}

        def one(path, total, out_bytes, launch, flush, note):
            lo, hi = shard.shard_range(total, rank, world)
            n = hi - lo
            nb_total = total * out_bytes
            gptr = L.b2g_dev_alloc(nb_total) if rank == 0 else None
            assert rank != 0 or gptr, "b2g_dev_alloc failed"
            handle = shard.broadcast_bytes(b.b2g_ipc_export(gptr) if rank == 0 else None, 64, device=dev)
            ipc_err = None
            try:
                peer = gptr if rank == 0 else b.b2g_ipc_open(handle)
            except Exception as e:           # no CUDA IPC in this container: report it, keep the other two variants
                peer, ipc_err = None, str(e)
            okf = torch.tensor([0 if peer is None else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            fused_ok = bool(okf.item())
            local = torch.zeros(n * out_bytes, dtype=torch.uint8, device=dev)
            gathered = torch.zeros(nb_total, dtype=torch.uint8, device=dev)
            t_c, _ = timed(lambda: launch(local.data_ptr(), lo, n), steps, warm, flush)
            t_f = None
            if fused_ok:
                t_f, _ = timed(lambda: launch(peer + lo * out_bytes, lo, n), steps, warm, flush)

            def with_nccl():
                launch(local.data_ptr(), lo, n)
                dist.all_gather_into_tensor(gathered, local)
            t_n, _ = timed(with_nccl, steps, warm, flush)
            # the gather alone (device time, max over ranks)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gathered, local)
            e1.record()
            torch.cuda.synchronize()
            t_g = shard.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device=dev)
            ident = None
            if rank == 0:
                full = torch.zeros(nb_total, dtype=torch.uint8, device=dev)
                launch(full.data_ptr(), 0, total)
                torch.cuda.synchronize()
                fused = torch.as_tensor(DevBuf(gptr, nb_total), device=dev)
                ident = {"fused_gather": bool(torch.equal(fused, full)) if fused_ok else None, "nccl_gather": bool(torch.equal(gathered, full)),
                         "sha_free_checksum": int(full[:: max(1, nb_total >> 16)].to(torch.int64).sum().item())}
                del full, fused
            barrier()
            if rank != 0 and peer is not None:
                b.b2g_ipc_close(peer)
            barrier()
            if rank == 0:
                L.b2g_dev_free(gptr)
            cfg = CFG[path]
            scale = (cfg["unit_bytes"] / 1e9) if cfg["unit"] == "GB/s" else 1.0
            val = lambda t: (total * steps / t * scale) if t else None            # noqa: E731
            one_gpu = results[path]["value"] / world if path in results else None
            r = {"units_total": total, "units_per_gpu": n, "unit": cfg["unit"], "compute": val(t_c), "fused_gather": val(t_f),
                 "nccl_gather": val(t_n), "ms": {"compute": 1e3 * t_c / steps, "fused_gather": 1e3 * t_f / steps if t_f else None,
                                                 "nccl_gather": 1e3 * t_n / steps, "nccl_allgather_alone": 1e3 * t_g},
                 "gathered_bytes": nb_total, "identical_to_single_gpu": ident, "one_gpu": one_gpu, "limiter": note}
            if one_gpu:
                r["vs_one_gpu"] = {k: (r[k] / one_gpu if r[k] else None) for k in ("compute", "fused_gather", "nccl_gather")}
            if ipc_err:
                r["ipc_error"] = ipc_err
            out[path] = r
            del local, gathered

        if "belt_ctr" in args.paths:
            secret = np.random.default_rng(10).integers(0, 256, 48, dtype=np.uint8).tobytes() if rank == 0 else None
            kiv = shard.broadcast_bytes(secret, 48, device=dev)
            st = b.BeltCTR(kiv[:32], kiv[32:])
            key, ctr = st.key_words, st.ctr_words
            one("belt_ctr", 1 << 26, 16,
                lambda dst, first, n: b.beltCTR_dev(dst, 0, n * 16, key, ctr, first, stream), False,
                "the gather: rank 0 can take 1 GiB no faster than one GPU's NVLink ingress, about as fast as one GPU "
                "computes it (BASELINE config 2 is a 1-GPU config)")
        if "bash512" in args.paths:
            g = torch.Generator(device=dev).manual_seed(1)
            msgs = torch.randint(0, 256, (1 << 20, 4096), dtype=torch.uint8, device=dev, generator=g)
            one("bash512", 1 << 20, 64,
                lambda dst, first, n: b.bashHashBatch_dev(dst, 256, msgs.data_ptr() + first * 4096, 4096, 4096, n, stream),
                False, "none expected: 2^20/N messages still fill every SM; digests are 1/64 of the input")
            del msgs
        if "bign_verify" in args.paths:
            units = 1 << 18
            rng = np.random.default_rng(2)
            priv = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            priv[:, 31] &= 0x7F
            hashes = rng.integers(0, 256, (units, 32), dtype=np.uint8)
            params = b.bignParamsStd()
            st1, pubs = b.bignPubkeyCalcBatch(params, priv)
            st2, sigs = b.bignSign2Batch(params, OID, hashes, priv)
            assert not st1.any() and not st2.any()
            bad = np.arange(0, units, 16)
            sigs[bad, bad % 48] ^= 1
            d_h, d_s, d_p = (torch.from_numpy(x).to(dev) for x in (hashes, sigs, pubs))
            one("bign_verify", units, 4,
                lambda dst, first, n: b.bignVerifyBatch_dev(dst, OID, d_h.data_ptr() + 32 * first, d_s.data_ptr() + 48 * first,
                                                            d_p.data_ptr() + 64 * first, n, stream),
                True, "wave quantisation: 2^18/N signatures = (1024/N) CTAs of 256 threads on 148 SMs x 3 CTA slots "
                      "(N=8: 128 CTAs, under one third of a wave, two warps per scheduler)")
            del d_h, d_s, d_p
        if "belt_ecb" in args.paths:
            units = 1 << 26
            g = torch.Generator(device=dev).manual_seed(3)
            keys = torch.randint(0, 256, (units, 32), dtype=torch.uint8, device=dev, generator=g)
            blocks = torch.randint(0, 256, (units, 16), dtype=torch.uint8, device=dev, generator=g)
            one("belt_ecb", units, 16,
                lambda dst, first, n: _chk(L.b2g_beltECBEncrBatch2_dev(dst, blocks.data_ptr() + 16 * first,
                                                                      keys.data_ptr() + 32 * first, n, stream)),
                False, "the gather of 1 GiB into one GPU (NVLink ingress), as for belt-CTR")
            del keys, blocks
        return out

    def _chk(code):
        assert code == 0, code

    strong = None
    if world > 1 and not args.no_strong:
        strong = strong_section()

    sampler.stop.set()
    clocks = sampler.summary()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # issue roofline: algorithmic integer work per second over the measured pipe peak
    mhz = clocks.get("sm_mhz")
    if "belt_ctr" in results:
        r = results["belt_ctr"]
        blocks_s = r["value"] * 1e9 / 16 / world
        r["issue_roofline"] = {"bound": "lds32 (224 conflict-free LDS.32 per block)", "achieved_Tops": blocks_s * 224 / 1e12,
                               "peak_Tops": issue.get("lds32"), "frac": blocks_s * 224 / 1e12 / issue["lds32"] if issue.get("lds32", 0) > 0 else None,
                               "alu_ops_per_block": 456, "alu_frac": blocks_s * 456 / 1e12 / issue["lop3"] if issue.get("lop3", 0) > 0 else None}
    if "belt_dwp" in results:
        r = results["belt_dwp"]
        blocks_s = r["value"] * 1e9 / 16 / world
        # LSU wavefronts per block: 224 (CTR LDS.32) + 32 LDS.128 x 4 (tag tables)
        r["issue_roofline"] = {"bound": "lsu wavefronts (224 LDS.32 + 32 LDS.128 x 4 per block)",
                               "achieved_Tops": blocks_s * 352 / 1e12, "peak_Tops": issue.get("lds32"),
                               "frac": blocks_s * 352 / 1e12 / issue["lds32"] if issue.get("lds32", 0) > 0 else None}
    if "belt_ecb" in results:
        r = results["belt_ecb"]
        blocks_s = r["value"] * 1e9 / 16 / world
        r["issue_roofline"] = {"bound": "lds32 (224 conflict-free LDS.32 per block)", "achieved_Tops": blocks_s * 224 / 1e12,
                               "peak_Tops": issue.get("lds32"), "frac": blocks_s * 224 / 1e12 / issue["lds32"] if issue.get("lds32", 0) > 0 else None}
    if "bash512" in results:
        r = results["bash512"]
        f_s = r["value"] * 1e9 / 4096 * 65 / world
        r["issue_roofline"] = {"bound": "alu (4272 LOP3/SHF per bash-f)", "achieved_Tops": f_s * 4272 / 1e12, "peak_Tops": issue.get("lop3"),
                               "frac": f_s * 4272 / 1e12 / issue["lop3"] if issue.get("lop3", 0) > 0 else None}
    if "bign_verify" in results:
        r = results["bign_verify"]
        v_s = r["value"] / world
        r["issue_roofline"] = {"bound": "imad.wide (2000 field mults x 72 wide multiply-adds per verify, SURVEY §8d)",
                               "achieved_Tops": v_s * 144000 / 1e12, "peak_Tops": issue.get("imad_wide"),
                               "frac": v_s * 144000 / 1e12 / issue["imad_wide"] if issue.get("imad_wide", 0) > 0 else None}
    if "bign_verify" in results and issue.get("imad_wide.cc chain", 0) > 0:
        # what the kernel really issues: carry-chained IMAD.WIDE.X runs at half the rate of the plain
        # form (microbench "imad_wide.cc chain"); executed count per verify from the committed ncu
        # source-page capture (profiles/README.md)
        v_s = results["bign_verify"]["value"] / world
        results["bign_verify"]["issue_roofline"]["carry_chain"] = {
            "executed_wide_mads_per_verify": NCU_BIGN_WIDE_PER_VERIFY, "achieved_Tops": v_s * NCU_BIGN_WIDE_PER_VERIFY / 1e12,
            "peak_Tops": issue["imad_wide.cc chain"], "frac": v_s * NCU_BIGN_WIDE_PER_VERIFY / 1e12 / issue["imad_wide.cc chain"]}
    if "bign_sign2" in results:
        r = results["bign_sign2"]
        v_s = r["value"] / world
        # 20 mixed additions (7M + 4S) from the fixed-base table + 2 log2(128) tree products + x = X/Z^2:
        # ~240 field products x 72 wide multiply-adds
        r["issue_roofline"] = {"bound": "imad.wide (240 field mults x 72 wide multiply-adds per signature)",
                               "achieved_Tops": v_s * 240 * 72 / 1e12, "peak_Tops": issue.get("imad_wide"),
                               "frac": v_s * 240 * 72 / 1e12 / issue["imad_wide"] if issue.get("imad_wide", 0) > 0 else None}
    if arm is not None:
        for path in args.paths:
            results[path]["cpu_baseline"] = arm.baseline(path)

    head = args.paths[0]
    h = results[head]
    cfg = CFG[head]
    if roof and h.get("e2e"):
        # the e2e roof of this box: raw concurrent pinned copies at this N (belt/bash e2e are copy-bound)
        for k, v in results.items():
            e = v.get("e2e")
            if not e:
                continue
            sec_per_step = world * CFG[k]["units"] * (CFG[k]["unit_bytes"] if CFG[k]["unit"] == "GB/s" else 1) / (e["value"] * (1e9 if CFG[k]["unit"] == "GB/s" else 1.0))
            floor = world * e["h2d_bytes_per_step"] / (roof["h2d"] * 1e9) + world * e["d2h_bytes_per_step"] / (roof["d2h"] * 1e9)
            e["copy_roof_GBs"] = roof
            e["copy_floor_ms"] = 1e3 * floor
            e["frac_of_copy_roof"] = floor / sec_per_step if sec_per_step > 0 else None
    also = {k: compact(v) for k, v in results.items() if k != head}
    config = {"workload": cfg["workload"], "l2": h["l2"], "units_per_gpu": cfg["units"],
              "sharding": "rank r takes units [r*U,(r+1)*U); key/iv/oid broadcast from rank 0 over NCCL; no data-path collective",
              "numa": numa, "also": also}
    if strong is not None:
        config["strong"] = {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "limiter"}
                            for k, v in strong.items()}
    line = {"metric": cfg["metric"], "value": h["value"], "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": config,
            "roofline": h["roofline"], "issue_roofline": h.get("issue_roofline"),
            "cpu_baseline": h.get("cpu_baseline"), "e2e": h.get("e2e"), "gather": h.get("gather"), "gpu_launches": h["gpu_launches"],
            "clocks": clocks, "issue_peaks_Tops": issue,
            # the architectural per-SM rates x 148 SMs x max SM clock, beside the measured ones
            "issue_peaks_arch_Tops": (lambda c: {"alu (lop3/shf/prmt/iadd) 64/clk/SM": 64 * 148 * c / 1e6,
                                                 "imad 64/clk/SM": 64 * 148 * c / 1e6, "lds32 32/clk/SM": 32 * 148 * c / 1e6,
                                                 "imad_wide plain 64/clk/SM, carry-chained 32/clk/SM": [64 * 148 * c / 1e6, 32 * 148 * c / 1e6]})(
                float(clocks.get("sm_max_mhz") or 1965)),
            "paths": {k: v for k, v in results.items() if k != head}, "strong": strong,
            "also": {k: compact(v) for k, v in results.items()}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
