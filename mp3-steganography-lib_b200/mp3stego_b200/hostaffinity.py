"""Host-side placement for one-process-per-GPU jobs: run the rank's threads -- and therefore, under Linux's first-touch policy,
place the pinned staging buffers it allocates afterwards -- on the NUMA node its GPU hangs off.  PCIe traffic of the host-buffer
(e2e) paths then stays on the socket that owns the root complex.  Reads /sys only; a no-op on single-node hosts (e.g. the
virtual machines of the GPU pool, which expose one node and numa_node = -1 for the device) and says so in its return value.

The reference has no counterpart (single process, no device)."""
import os
from typing import Dict, List, Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s: str) -> List[int]:
    out: List[int] = []
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def numa_nodes() -> Dict[int, List[int]]:
    """{node: cpus} from /sys/devices/system/node."""
    base = "/sys/devices/system/node"
    nodes = {}
    try:
        names = os.listdir(base)
    except OSError:
        return nodes
    for n in names:
        if n.startswith("node") and n[4:].isdigit():
            cl = _read(os.path.join(base, n, "cpulist"))
            if cl:
                nodes[int(n[4:])] = _parse_cpulist(cl)
    return nodes


def gpu_numa_node(pci_bus_id: str) -> int:
    """NUMA node of a PCI device ('0000:9c:00.0' or nvidia-smi's '00000000:9C:00.0'); -1 when the platform does not say."""
    bdf = pci_bus_id.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    v = _read(f"/sys/bus/pci/devices/{bdf}/numa_node")
    try:
        return int(v) if v is not None else -1
    except ValueError:
        return -1


def bind_to_device(pci_bus_id: str, local_rank: int = 0, local_world: int = 1) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (sched_setaffinity) BEFORE it allocates pinned memory.  When the
    platform gives the device no node but the host has several, the ranks are spread round-robin over the nodes so that at least
    their staging buffers do not pile up on one socket.  Returns what was done (goes into the bench line)."""
    nodes = numa_nodes()
    info = dict(numa_nodes=len(nodes), gpu_node=gpu_numa_node(pci_bus_id), bound=False, cpus=len(os.sched_getaffinity(0)))
    if len(nodes) <= 1:
        info["note"] = "single NUMA node exposed: nothing to bind"
        return info
    node = info["gpu_node"]
    if node < 0 or node not in nodes:
        node = sorted(nodes)[local_rank * len(nodes) // max(local_world, 1) % len(nodes)]
        info["note"] = "device reports no NUMA node: ranks spread over the nodes"
    allowed = sorted(set(nodes[node]) & os.sched_getaffinity(0))
    if not allowed:
        info["note"] = "node CPUs are outside this process's cpuset"
        return info
    os.sched_setaffinity(0, allowed)
    info.update(bound=True, node=node, cpus=len(allowed))
    return info
