"""Multi-GPU sharding of the importance-sampling integral.

Events are independent, so the global event range [0, N_total) is split contiguously, one process per
GPU; the Philox counter is the GLOBAL event index, so the estimator's inputs do not depend on the number
of ranks. The only collective is one all-reduce of the 3 accumulators (sum w, sum w^2, n_valid) at the end
-- never per chunk. The reference itself has no explicit collective (DDP only).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous [begin, end) owned by `rank`; the union over ranks is exactly [0, n_total)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return (n_total * rank) // world_size, (n_total * (rank + 1)) // world_size


def chunk_ranges(begin, end, chunk):
    """Split [begin, end) into consecutive chunks of at most `chunk` events."""
    out = []
    b = begin
    while b < end:
        e = min(b + chunk, end)
        out.append((b, e))
        b = e
    return out


def allreduce_accumulators(acc, group=None):
    """In-place SUM all-reduce of the double[3] accumulator (NCCL on GPUs, gloo in the CPU tests).
    n_valid travels as a double: exact up to 2^53 events."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


class ShardedIntegrator:
    """Runs an `ImportanceSampler` over this rank's shard of [0, n_total) in chunks and reduces once.

    `phi_provider(begin, end)` returns the float32 [T, end-begin, F, 3K-1] spline parameters for the
    global events [begin, end) (in MEMFlow: the flow's hyper-network output; in the benchmark: a resident
    synthetic buffer)."""

    def __init__(self, sampler, n_total, chunk, rank=0, world_size=1, group=None):
        self.sampler = sampler
        self.n_total = int(n_total)
        self.chunk = int(chunk)
        self.rank, self.world_size, self.group = rank, world_size, group
        self.begin, self.end = shard_range(self.n_total, rank, world_size)

    def run(self, phi_provider, acc=None, state_path=None, save_every=0):
        """Device-resident parameters. With `state_path` the run is restartable: every `save_every` chunks (and at the
        end) the shard's accumulators and the index of the next chunk are written (one host sync per save); a later call
        with the same path resumes after the last saved chunk and gives bit-identical accumulators (the Philox counter
        is the global event index, the reduction is deterministic)."""
        from .estimator import new_accumulator
        if acc is None:
            acc = new_accumulator(self.sampler.device)
        chunks = chunk_ranges(self.begin, self.end, self.chunk)
        start = self._resume(state_path, acc)
        for i in range(start, len(chunks)):
            b, e = chunks[i]
            self.sampler.run_chunk(phi_provider(b, e), event_offset=b, acc=acc)
            if state_path and save_every and (i + 1) % save_every == 0:
                self._save(state_path, acc, i + 1)
        if state_path:
            self._save(state_path, acc, len(chunks))
        return allreduce_accumulators(acc.clone() if state_path else acc, self.group)

    def run_from_host(self, host_phi_provider, acc=None, n_copy_streams=2, pipeline=None):
        """Host-resident parameters: `host_phi_provider(begin, end)` returns a (preferably pinned) CPU tensor
        [T, end-begin, F, 3K-1] for the global events [begin, end). Copies are double-buffered against the kernels on
        `n_copy_streams` copy streams (HostPipeline); one all-reduce at the end. Returns (acc, pipeline.stats())."""
        from .estimator import new_accumulator
        if acc is None:
            acc = new_accumulator(self.sampler.device)
        chunks = chunk_ranges(self.begin, self.end, self.chunk)
        if not chunks:
            return allreduce_accumulators(acc, self.group), {}
        for b, e in chunks:
            hp = host_phi_provider(b, e)
            if pipeline is None:
                pipeline = HostPipeline(self.sampler, (hp.shape[0], self.chunk, hp.shape[2], hp.shape[3]), dtype=hp.dtype,
                                        n_copy_streams=n_copy_streams)
            pipeline.submit(hp, b, acc)
        stats = pipeline.stats()
        return allreduce_accumulators(acc, self.group), stats

    # ---- restartable accumulator (checkpoint row): sum w, sum w^2, n_valid + the next chunk index ----
    def _state(self):
        return {"n_total": self.n_total, "chunk": self.chunk, "rank": self.rank, "world_size": self.world_size,
                "seed": self.sampler.seed, "begin": self.begin, "end": self.end}

    def _save(self, path, acc, next_chunk):
        import json, os, struct
        a = acc.detach().cpu().tolist()
        # the accumulators are stored as IEEE-754 bit patterns: a resumed run continues from exactly the same doubles
        rec = dict(self._state(), next_chunk=int(next_chunk), acc_hex=[struct.pack(">d", v).hex() for v in a])
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(rec, f)
        os.replace(tmp, path)

    def _resume(self, path, acc):
        import json, os, struct
        if not path or not os.path.exists(path):
            return 0
        rec = json.load(open(path))
        for k, v in self._state().items():
            if rec.get(k) != v:
                raise ValueError("checkpoint %s was written for %s = %r, this run has %r" % (path, k, rec.get(k), v))
        vals = [struct.unpack(">d", bytes.fromhex(h))[0] for h in rec["acc_hex"]]
        acc.copy_(torch.tensor(vals, dtype=torch.float64))
        return int(rec["next_chunk"])


# ---------------------------------------------------------------------------------------------------------------------
# Host-resident parameters: the reference-facing call when the flow's hyper-network output (or a file of stored
# parameters) lives in host memory. Moved here from bench.py (round 1) so that the timed "plugin call with host buffers"
# is a product API.
# ---------------------------------------------------------------------------------------------------------------------
def gpu_cpu_affinity(device_index):
    """(numa_node, cpu list) of the PCIe root the GPU hangs off, from sysfs; (None, None) when it cannot be read."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = []
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.extend(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        return node, cpus
    except Exception:
        return None, None


def bind_to_gpu_cpus(device_index):
    """Pin this process to the CPUs local to its GPU (so that pinned staging buffers allocated afterwards are first
    touched on that NUMA node). Returns the (numa_node, n_cpus) it bound to, or (None, 0)."""
    import os
    node, cpus = gpu_cpu_affinity(device_index)
    if cpus:
        try:
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0))) or cpus
            os.sched_setaffinity(0, allowed)
            return node, len(allowed)
        except Exception:
            pass
    return None, 0


class HostPipeline:
    """Streams host-resident spline parameters through a sampler: chunk i + 1 is copied host -> device on `n_copy_streams`
    copy streams (the chunk is split along its leading transform / event dimension, one slice per stream) while chunk i
    runs on the compute stream; two device buffers, events both ways. No host synchronisation inside the loop.

    stats(): bytes copied, the copy streams' busy time measured with CUDA events, achieved H2D GB/s.
    """

    def __init__(self, sampler, chunk_shape, dtype=torch.float32, n_copy_streams=2, device=None):
        self.sampler = sampler
        self.device = torch.device(device if device is not None else sampler.device)
        self.shape = tuple(int(s) for s in chunk_shape)           # [T, N, F, 3K-1]
        self.bufs = [torch.empty(self.shape, dtype=dtype, device=self.device) for _ in range(2)]
        self.n_copy = max(1, int(n_copy_streams))
        self.copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_copy)]
        self.copied = [[torch.cuda.Event() for _ in range(self.n_copy)] for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.t0 = [torch.cuda.Event(enable_timing=True) for _ in range(self.n_copy)]
        self.t1 = [torch.cuda.Event(enable_timing=True) for _ in range(self.n_copy)]
        self.bytes = 0
        self._i = 0
        # scratch of the chain's two launches (u, log q), owned here: no stream-ordered allocation per chunk
        F = self.shape[2]
        self._u = torch.empty((self.shape[1], F), dtype=torch.float64, device=self.device)
        self._lq = torch.empty(self.shape[1], dtype=torch.float64, device=self.device)

    def pinned_chunk(self):
        """A pinned host staging tensor of the chunk shape (allocate after bind_to_gpu_cpus)."""
        return torch.empty(self.shape, dtype=self.bufs[0].dtype).pin_memory()

    def _pieces(self, T, n_events):
        """Contiguous pieces (t, lo, hi) of a [T, n, F, 3K-1] chunk, dealt round-robin to the copy streams. Every piece is
        a contiguous block of host memory: a strided H2D copy would be staged through a pageable bounce buffer by torch
        (6 GB/s instead of 55)."""
        parts = max(1, -(-self.n_copy // T))          # pieces per transform so that there are >= n_copy pieces
        per = -(-n_events // parts)
        out = []
        for t in range(T):
            for j in range(parts):
                lo, hi = j * per, min(n_events, (j + 1) * per)
                if lo < hi:
                    out.append((t, lo, hi))
        return out

    def submit(self, host_phi, event_offset, acc):
        """Queue one chunk: async H2D of host_phi [T, n, F, 3K-1] (n <= chunk events; pinned memory for a truly asynchronous
        copy), then the chain on it, accumulating into acc. Returns immediately."""
        b = self._i & 1
        first = self._i < 2
        main = torch.cuda.current_stream(self.device)
        n = int(host_phi.shape[1])
        dst = self.bufs[b][:, :n]
        used = set()
        for j, (t, lo, hi) in enumerate(self._pieces(int(host_phi.shape[0]), n)):
            s = j % self.n_copy
            cs = self.copy_streams[s]
            with torch.cuda.stream(cs):
                if s not in used:
                    used.add(s)
                    if not first:
                        cs.wait_event(self.consumed[b])
                    if self._i == 0:
                        self.t0[s].record(cs)
                dst[t, lo:hi].copy_(host_phi[t, lo:hi], non_blocking=True)
        for s in used:
            cs = self.copy_streams[s]
            with torch.cuda.stream(cs):
                self.copied[b][s].record(cs)
                self.t1[s].record(cs)
            main.wait_event(self.copied[b][s])
        self.bytes += host_phi.numel() * host_phi.element_size()
        phi = dst if n == self.shape[1] else dst.contiguous()
        u, lq = self.sampler.sample_flow(phi, event_offset=event_offset, u=self._u[:n], logq=self._lq[:n])
        self.sampler.weigh(u, lq, acc=acc)
        self.consumed[b].record(main)
        self._i += 1
        return acc

    def stats(self):
        torch.cuda.synchronize(self.device)
        ms = 0.0
        if self._i:
            for a, b in zip(self.t0, self.t1):
                try:
                    ms = max(ms, a.elapsed_time(b))
                except Exception:   # a copy stream that never got a piece
                    pass
        return {"h2d_bytes": int(self.bytes), "copy_window_ms": ms, "h2d_GBps": (self.bytes / ms / 1e6) if ms > 0 else None,
                "copy_streams": self.n_copy, "chunks": self._i}
