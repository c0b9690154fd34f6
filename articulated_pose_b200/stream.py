"""Mixed-category cloud streams (BASELINE.json configs[4]: all five categories, sharded over the GPUs of one box).

The reference runs one `main.py --test` + `pose_multi_process.py` per category because every category has its own
checkpoint, part count K and cloud size N (global_info.py, lib/dataset.py:35).  A launch of the kernels here is
homogeneous in (K, N, weights) too, so a mixed stream is bucketed by category: each rank takes its contiguous slice of
the stream (dist.shard_range -- the pose_multi_process.py:54-63 rule), groups the slice by category, cuts every group
into batches and pushes them through that category's AncshPipeline (the pipelines of different categories share the
device; their batches are pipelined through AncshPipeline.run_many).  Results come back in stream order; the one
collective is the final all-gather of the per-cloud pose records padded to the widest category
(dist.pack_records / gather_records).
"""
import numpy as np

from . import dist as adist


def bucket_by_category(items):
    """items: sequence of (category, payload) -> {category: [positions in the sequence]} (order preserved)."""
    buckets = {}
    for pos, (cat, _) in enumerate(items):
        buckets.setdefault(cat, []).append(pos)
    return buckets


def batches(positions, batch):
    """Cut a bucket into consecutive batches of at most `batch` positions."""
    return [positions[i:i + batch] for i in range(0, len(positions), batch)]


def record_matrix(results, parts, k_max):
    """Per-cloud pose records of clouds with different part counts as ONE float64 matrix for the all-gather: row =
    [K, record of dist.pack_records zero-padded to the width of k_max parts]."""
    width = 1 + adist.record_width(k_max)
    out = np.zeros((len(results), width), np.float64)
    for i, (r, K) in enumerate(zip(results, parts)):
        out[i, 0] = K
        rec = adist.pack_records([r], K)[0]
        out[i, 1:1 + rec.shape[0]] = rec
    return out


class MixedStream:
    """pipelines: {category: AncshPipeline}; `load(category, payload)` -> (P (N,3) float32, joint_cls (N,) int32), or
    `load_batch(category, [payloads])` -> (P (n,N,3), joint_cls (n,N)) when the caller can fetch a whole batch at once."""

    def __init__(self, pipelines, load=None, batch=256, rank=0, world=1, pad=True, load_batch=None):
        if load is None and load_batch is None:
            raise ValueError("MixedStream needs load or load_batch")
        self.pipelines, self.load, self.load_batch = pipelines, load, load_batch
        self.batch, self.rank, self.world, self.pad = int(batch), int(rank), int(world), bool(pad)

    def my_slice(self, n_items):
        return adist.shard_range(n_items, self.rank, self.world)

    def _load_chunk(self, cat, payloads):
        if self.load_batch is not None:
            P, jc = self.load_batch(cat, payloads)
            P, jc = np.asarray(P, np.float32), np.asarray(jc, np.int32)
        else:
            loaded = [self.load(cat, p) for p in payloads]
            P = np.stack([l[0] for l in loaded]).astype(np.float32)
            jc = np.stack([l[1] for l in loaded]).astype(np.int32)
        # a ragged last batch is padded with copies of its last cloud (results dropped by the caller): every launch of a
        # category then has ONE shape, so the pipelined path never allocates pinned / device buffers mid-stream
        pad = self.batch - P.shape[0] if self.pad else 0
        if pad > 0:
            P = np.concatenate([P, np.repeat(P[-1:], pad, 0)])
            jc = np.concatenate([jc, np.repeat(jc[-1:], pad, 0)])
        return P, jc

    def run(self, items, unpack=True, records_k_max=None):
        """items: the WHOLE stream [(category, payload)] (the same list on every rank).  Returns (start, end, results):
        this rank's slice bounds and its per-cloud results in stream order -- result dicts (unpack=True: pose.unpack_results
        entries, else per-cloud slices of the pose arrays), or, with records_k_max, ONE float64 matrix (n, 1 + width(k_max))
        of gather-ready records (dist.records_from_arrays) built without any per-cloud Python work."""
        s, e = self.my_slice(len(items))
        mine = list(items[s:e])
        as_records = records_k_max is not None
        results = [None] * len(mine)
        rec = np.zeros((len(mine), 1 + adist.record_width(records_k_max)), np.float64) if as_records else None
        # enqueue every category before collecting any: the forwards of the next category overlap the pose tail of the
        # previous one (each pipeline has its own slots and pose streams)
        started = []
        for cat, positions in bucket_by_category(mine).items():
            pipe = self.pipelines[cat]
            chunks = batches(positions, self.batch)
            work = [self._load_chunk(cat, [mine[p][1] for p in chunk]) for chunk in chunks]
            full = [w for w in work if w[0].shape[0] == work[0][0].shape[0]]
            begin = getattr(pipe, "run_many_begin", None)       # any object with run_many() works; the split form overlaps
            finish = begin(full, unpack=unpack and not as_records) if begin else \
                (lambda r=pipe.run_many(full, unpack=unpack and not as_records): r)
            started.append((pipe, chunks, work, full, finish))
        for pipe, chunks, work, full, finish in started:
            outs = finish()
            for w in work[len(full):]:
                outs += pipe.run_many([w], unpack=unpack and not as_records)
            for chunk, out in zip(chunks, outs):
                if as_records:
                    rec[np.asarray(chunk)] = adist.records_from_arrays(out, pipe.K, records_k_max)[:len(chunk)]
                elif unpack:
                    for p, r in zip(chunk, out):
                        results[p] = r
                else:
                    for i, p in enumerate(chunk):
                        results[p] = {k: v[i] for k, v in out.items()}
        return s, e, (rec if as_records else results)

    def gather(self, results, parts, k_max, device=None):
        """The path's single collective: every rank's records -> the full (n_clouds, 1 + width(k_max)) matrix."""
        return adist.gather_records(record_matrix(results, parts, k_max), device=device)
