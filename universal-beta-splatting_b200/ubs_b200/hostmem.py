"""Host-side placement for the end-to-end path: pinned image buffers should live on the NUMA node the GPU hangs off.

cudaHostAlloc places pages on the node of the calling thread (first touch under the default policy), so binding the
process to the GPU's node BEFORE the pinned buffers are allocated is all it takes.  One process per GPU (torchrun), so
the binding is per process.  Everything here is best effort: on a single-node host, or without sysfs / NVML, it does
nothing and says so."""
import os
from typing import Optional


def gpu_numa_node(physical_index: int) -> Optional[int]:
    """NUMA node of GPU `physical_index` (NVML order) from sysfs, or None when unknown / single node."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        node = int(open(path).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node: int):
    def parse(s):
        out = []
        for part in s.strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                out.extend(range(int(a), int(b) + 1))
            elif part:
                out.append(int(part))
        return out

    return parse(open("/sys/devices/system/node/node%d/cpulist" % node).read())


def bind_to_gpu_node(physical_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for the bench line)."""
    info = {"gpu": physical_index, "numa_node": None, "bound": False}
    try:
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        n_nodes = 1
    info["numa_nodes"] = n_nodes
    node = gpu_numa_node(physical_index)
    info["numa_node"] = node
    if node is None or n_nodes < 2:
        return info
    try:
        allowed = os.sched_getaffinity(0)
        cpus = set(node_cpus(node)) & allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"], info["cpus"] = True, len(cpus)
    except Exception as e:  # pragma: no cover
        info["error"] = str(e)
    return info
