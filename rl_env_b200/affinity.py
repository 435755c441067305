"""Host-side placement for the host-buffer (end-to-end) path: run a rank's thread and put its pinned
buffers on the NUMA node its GPU hangs off.

With one process per GPU all doing 57 MB device-to-host copies per step, ranks that share the default
affinity (all CPUs, first-touch memory on whatever node the thread happens to run on) push most of the
traffic through one memory controller and the inter-socket link (SCALE_r01: 12 GB/s per GPU at 8 ranks
against 49 GB/s alone).  `bind_to_gpu` narrows the CPU affinity to the GPU's node (within the CPUs the
process is allowed to use) and sets the memory policy to prefer that node, BEFORE the pinned buffers are
allocated.  Everything is best effort: on any failure the process is left as it was."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of CUDA device `device_index` (sysfs numa_node of its PCI function), None if unknown."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
    except Exception:
        return None
    text = _read(f"/sys/bus/pci/devices/{bdf}/numa_node")
    if text is None:
        return None
    node = int(text)
    return node if node >= 0 else None


def bind_to_gpu(device_index: int) -> Dict[str, object]:
    """Bind the calling process to the NUMA node of its GPU; returns what was done (for the bench line)."""
    info: Dict[str, object] = {"numa_node": None, "cpus": None, "mempolicy": False}
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    info["numa_node"] = node
    cpulist = _read(f"/sys/devices/system/node/node{node}/cpulist")
    try:
        allowed = os.sched_getaffinity(0)
        want = set(_parse_cpulist(cpulist)) & allowed if cpulist else set()
        if want:
            os.sched_setaffinity(0, want)
            info["cpus"] = len(want)
    except (OSError, AttributeError, ValueError):
        pass
    try:                                   # set_mempolicy(MPOL_PREFERRED, {node}): pinned buffers allocated from now on land there
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] = 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(16 * 8 * ctypes.sizeof(ctypes.c_ulong)))
        info["mempolicy"] = rc == 0
    except Exception:
        pass
    return info
