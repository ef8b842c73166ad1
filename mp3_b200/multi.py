"""Sharding streams across GPUs: the path has no exchange step (streams never interact), so
multi-GPU is N independent contexts and a partition of the stream list.  No collectives.
"""
import threading

import numpy as np


def partition(sizes, nparts):
    """Split stream indices 0..n-1 into `nparts` contiguous ranges of near-equal total size.

    `sizes` is the per-stream work estimate (bytes or frames).  Returns a list of (lo, hi) index
    ranges; ranges are contiguous so that each rank's PCM stays in input order."""
    sizes = np.asarray(sizes, np.float64)
    n = sizes.size
    if nparts <= 0:
        raise ValueError("nparts must be positive")
    cum = np.concatenate([[0.0], np.cumsum(sizes)])
    total = cum[-1]
    bounds = [0]
    for p in range(1, nparts):
        target = total * p / nparts
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(nparts)]


def merge(parts, ranges, n):
    """Inverse of partition for per-stream results: parts[r] is the list for range ranges[r]."""
    out = [None] * n
    for part, (lo, hi) in zip(parts, ranges):
        assert len(part) == hi - lo
        out[lo:hi] = part
    return out


def _parse_cpulist(txt):
    cpus = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that the pinned
    staging buffers it allocates afterwards (first touch) are local to that GPU's PCIe root: with one
    process per GPU, cross-socket staging is what caps the aggregate host <-> device bandwidth.
    Returns a short description, or None when the topology cannot be read (then nothing changes)."""
    import os
    try:
        import torch
        bdf = torch.cuda.get_device_properties(device).pci_bus_id  # not on every torch build
    except Exception:
        bdf = None
    try:
        if bdf is None:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            if isinstance(bdf, bytes):
                bdf = bdf.decode()
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return "numa node %d (%d cpus)" % (node, len(allowed))
    except Exception:
        return None


class MultiDecoder:
    """One Decoder (context) per device, each driven by its own thread."""

    def __init__(self, devices, **kw):
        import mp3_b200
        self.decs = [mp3_b200.Decoder(device=d, **kw) for d in devices]

    def decode(self, streams):
        """Decode a list of byte strings; returns per-stream PCM arrays [samples, channels]."""
        ranges = partition([len(s) for s in streams], len(self.decs))
        parts = [None] * len(self.decs)

        def work(r):
            lo, hi = ranges[r]
            dec = self.decs[r]
            dec.decode_batch(streams[lo:hi])
            arena = dec.fetch_pcm()
            parts[r] = [dec.stream_pcm(i, arena).copy() for i in range(hi - lo)]

        th = [threading.Thread(target=work, args=(r,)) for r in range(len(self.decs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return merge(parts, ranges, len(streams))

    def close(self):
        for d in self.decs:
            d.close()
