"""SNP sharding across ranks: the analysed SNP list is cut into contiguous
ranges (keeps output order == file order), every rank runs the whole per-SNP path on its range with no
data-path communication, and ONE gather of the SUMSTAT rows (64 B per SNP) ends the run."""
import numpy as np

from .api import SUMSTAT_DTYPE


def snp_range(n_snps, rank, world):
    """Contiguous range [lo, hi) of rank's SNPs; sizes differ by at most one."""
    base, rem = divmod(n_snps, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_sumstat(local, n_snps, group=None, dst=0):
    """Gather the per-rank SUMSTAT arrays (numpy structured, len == range size) on `dst` in SNP order.
    Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [snp_range(n_snps, r, world)[1] - snp_range(n_snps, r, world)[0] for r in range(world)]
    mx = max(sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros((mx, 8), dtype=torch.float64, device=dev)
    if len(local):
        buf[:len(local)] = torch.from_numpy(np.ascontiguousarray(local).view(np.float64).reshape(-1, 8)).to(dev)
    out = torch.empty((world * mx, 8), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, mx, 8)
    if rank != dst:
        return None
    parts = [out[r, :sizes[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts, axis=0).reshape(-1).view(SUMSTAT_DTYPE)


def combine_partial_kinship(K_local, ns_local, group=None):
    """-gk across ranks: every rank accumulated K over ITS SNP range and scaled it by its own
    1/ns (gb200_kin_finish); the global matrix is sum_r ns_r K_r / sum_r ns_r -- one all-reduce of n^2 doubles plus one
    of the SNP counts.  K_local: torch tensor (CUDA with NCCL, CPU with gloo), modified in place and returned."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([float(ns_local)], dtype=torch.float64, device=K_local.device)
    if ns_local == 0:
        K_local.zero_()                      # a rank without SNPs contributes nothing (and never a NaN * 0)
    else:
        K_local.mul_(float(ns_local))
    dist.all_reduce(K_local, group=group)
    dist.all_reduce(cnt, group=group)
    K_local.div_(cnt.item())
    return K_local, int(cnt.item())


class _DevArray:
    """__cuda_array_interface__ view of a device buffer owned by the library (no copy)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape):
    import torch
    return torch.as_tensor(_DevArray(ptr, shape), device=torch.device("cuda", torch.cuda.current_device()))
