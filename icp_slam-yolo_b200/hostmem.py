"""Pinned host buffers placed on the NUMA node of the GPU that will read them.

The end-to-end path (`registration.HostPipeline`) is bounded by the host-to-device copy once several
GPUs of a node copy at the same time: every GPU then pulls from host DRAM, and a buffer
that sits on the other socket crosses the inter-socket link first.  `near_gpu(device)` binds the calling
thread (CPU affinity + preferred memory node) to the GPU's node for the duration of the `with` block, so
that `cudaHostAlloc` / first touch inside it lands there; `pin_near_gpu` does that for one array.

Linux only; every step degrades to a no-op (and says so in the returned record) when the platform does
not expose the topology — a VM without NUMA information reports node -1 for every PCI device.
"""
import contextlib
import ctypes
import os

import torch

_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_SYS_SET_MEMPOLICY = 238            # x86_64


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11}."""
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_pci_address(device_index):
    p = torch.cuda.get_device_properties(device_index)
    return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)


def gpu_numa_node(device_index, sysfs="/sys"):
    """NUMA node of the GPU's PCI function, or None when the platform does not say."""
    text = _read(os.path.join(sysfs, "bus/pci/devices", gpu_pci_address(device_index), "numa_node"))
    try:
        node = int(text)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node, sysfs="/sys"):
    return parse_cpulist(_read(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)))


def host_nodes(sysfs="/sys"):
    return sorted(parse_cpulist(_read(os.path.join(sysfs, "devices/system/node/online"))))


def _set_mempolicy(mode, node):
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        if mode == _MPOL_DEFAULT:
            return libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_DEFAULT, None, 0) == 0
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(_SYS_SET_MEMPOLICY, mode, mask, 16 * 64 + 1) == 0
    except Exception:
        return False


@contextlib.contextmanager
def near_gpu(device_index, sysfs="/sys"):
    """Bind this thread to the GPU's NUMA node while host buffers are allocated; yields a record
    {"gpu", "pci", "node", "host_nodes", "bound", "why"}."""
    rec = {"gpu": int(device_index), "pci": None, "node": None, "host_nodes": host_nodes(sysfs), "bound": False,
           "why": None}
    try:
        rec["pci"] = gpu_pci_address(device_index)
        rec["node"] = gpu_numa_node(device_index, sysfs)
    except Exception as e:                                    # no CUDA device: nothing to bind to
        rec["why"] = "no device properties: %s" % e
    old_aff = None
    policy = False
    if rec["node"] is None:
        rec["why"] = rec["why"] or "the platform reports no NUMA node for this PCI device"
    elif len(rec["host_nodes"]) < 2:
        rec["why"] = "single NUMA node"
    else:
        cpus = node_cpus(rec["node"], sysfs) & os.sched_getaffinity(0)
        if cpus:
            old_aff = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
        policy = _set_mempolicy(_MPOL_PREFERRED, rec["node"])
        rec["bound"] = bool(cpus) or policy
        if not rec["bound"]:
            rec["why"] = "neither CPU affinity nor memory policy could be set"
    try:
        yield rec
    finally:
        if policy:
            _set_mempolicy(_MPOL_DEFAULT, 0)
        if old_aff is not None:
            os.sched_setaffinity(0, old_aff)


def pin_near_gpu(array, device_index):
    """Pinned copy of a NumPy array / CPU tensor, allocated on the GPU's NUMA node when known.
    Returns (tensor, record)."""
    t = array if isinstance(array, torch.Tensor) else torch.from_numpy(array)
    with near_gpu(device_index) as rec:
        pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned.copy_(t)
    return pinned, rec
