"""Device-pointer entry points for callers that keep their arrays in HBM as
torch CUDA tensors (bench.py, the rank-mode tests).  torch is plumbing here:
it owns the allocations and the stream; all arithmetic is in liblpmgpu.so."""
import ctypes as C

import torch

from ._lib import lib, check


def _p(t, dtype=torch.float64):
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.device, t.dtype, t.is_contiguous())
    return C.c_void_p(t.data_ptr())


def _stream(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def bve_velocity_dev(x, y, z, relvort, area, mask, radius, ibeg, iend, u, v, w, stream=None):
    check(lib.lpm_bve_velocity_dev(x.numel(), _p(x), _p(y), _p(z), _p(relvort), _p(area), _p(mask, torch.int32),
                                   float(radius), int(ibeg), int(iend), _p(u), _p(v), _p(w), _stream(stream)))


def bve_stream_dev(x, y, z, relvort, absvort, area, mask, radius, ibeg, iend, rs, as_, stream=None):
    check(lib.lpm_bve_stream_dev(x.numel(), _p(x), _p(y), _p(z), _p(relvort), _p(absvort), _p(area),
                                 _p(mask, torch.int32), float(radius), int(ibeg), int(iend), _p(rs), _p(as_),
                                 _stream(stream)))


def plane_velocity_dev(x, y, vort, area, mask, ibeg, iend, u, v, stream=None):
    check(lib.lpm_plane_velocity_dev(x.numel(), _p(x), _p(y), _p(vort), _p(area), _p(mask, torch.int32),
                                     int(ibeg), int(iend), _p(u), _p(v), _stream(stream)))


def betaplane_velocity_dev(x, y, relvort, area, mask, ibeg, iend, u, v, stream=None):
    check(lib.lpm_betaplane_velocity_dev(x.numel(), _p(x), _p(y), _p(relvort), _p(area), _p(mask, torch.int32),
                                         int(ibeg), int(iend), _p(u), _p(v), _stream(stream)))


def pse_laplacian_sphere_dev(x, y, z, f, area, mask, eps, sphere_radius, ibeg, iend, lap, stream=None):
    check(lib.lpm_pse_laplacian_sphere_dev(x.numel(), _p(x), _p(y), _p(z), _p(f), _p(area), _p(mask, torch.int32),
                                           float(eps), float(sphere_radius), int(ibeg), int(iend), _p(lap),
                                           _stream(stream)))


class _SharedArray:
    """__cuda_array_interface__ view of library-owned device memory (a shared slab)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def shared_tensors(count, n, device):
    """COLLECTIVE (rank mode): `count` float64 tensors of n entries carved from ONE shared slab
    (lpm_comm_alloc_shared).  Direct sums whose outputs are these tensors deliver every rank's
    slice to every rank by NVLink peer stores: no allgather_slices_dev afterwards.  Returns
    (tensors, slab address); free with api.comm_free_shared(address) once the tensors are dropped."""
    from . import api
    stride = (n * 8 + 255) // 256 * 256
    base = api.comm_alloc_shared(stride * count)
    ts = [torch.as_tensor(_SharedArray(base + k * stride, (n,)), device=device) for k in range(count)]
    for k, t in enumerate(ts):
        assert t.data_ptr() == base + k * stride
        t.zero_()
    return ts, base


def allgather_slices_dev(tensors, stream=None):
    """The reference's MPI_BCAST loop (src/SphereBVESolver.f90:422-429) over NCCL, in place."""
    n = tensors[0].numel()
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    check(lib.lpm_comm_allgather_slices_dev(len(tensors), arr, n, _stream(stream)))
