!> @file lpm_gpu.f90
!> ISO_C_BINDING interface to liblpmgpu.so (include/lpm_gpu.h): the reference-side
!> binding a maintainer adds to lpm-v2's src/ (together with lpm_gpu_interface.inc, and to the source list in
!> src/CMakeLists.txt:1-16) to route the O(N^2) direct sums to the B200 library.
!> NOT compiled in this repository's image (no Fortran compiler is installed);
!> see INTEGRATION.md for the patched bodies of the private kernels that call it.
module LpmGpuModule
use iso_c_binding
use NumberKindsModule
use LoggerModule
implicit none
public
!> Every entry point of include/lpm_gpu.h, generated from the header by tools/gen_fortran_interface.py
!> (scalars by value, arrays by reference, handles / device addresses / streams as type(c_ptr)):
!> the host sums (lpm_bve_velocity ... lpm_swe_sphere_rhs_integrals), the device-pointer sums (*_dev), the
!> resident solvers (lpm_bve_solver_*, lpm_plane_solver_*, lpm_betaplane_solver_*), rank mode
!> (lpm_gpu_init_rank, lpm_comm_*), pinning, LoadBalance and the measurement hooks.
interface
include "lpm_gpu_interface.inc"
end interface

contains

!> Call once after MPI_INIT.  With the GPU library the driver runs as ONE rank
!> (mpirun -np 1): numProcs = 1, so every MPISetup slice is 1..N and the
!> residual MPI_BCAST-to-self loops in untouched code are no-ops.
subroutine LpmGpuInit(aLog, nDevices)
	type(Logger), intent(inout) :: aLog
	integer(kint), intent(out) :: nDevices
	integer(c_int) :: ierr, nUsed
	ierr = lpm_gpu_init(0_c_int, nUsed)
	nDevices = nUsed
	call LpmGpuCheck(aLog, ierr, "lpm_gpu_init")
end subroutine

subroutine LpmGpuFinalize()
	integer(c_int) :: ierr
	ierr = lpm_gpu_finalize()
end subroutine

!> The reference never aborts on this path: errors are logged at ERROR level,
!> which clears the global testPass (Logger.f90:158-160).
subroutine LpmGpuCheck(aLog, ierr, where)
	type(Logger), intent(inout) :: aLog
	integer(c_int), intent(in) :: ierr
	character(len=*), intent(in) :: where
	character(kind=c_char), pointer :: cmsg(:)
	character(len=512) :: msg
	integer :: i
	if (ierr == 0) return
	call c_f_pointer(lpm_gpu_last_error(), cmsg, [512])
	msg = ""
	do i = 1, 512
		if (cmsg(i) == c_null_char) exit
		msg(i:i) = cmsg(i)
	enddo
	call LogMessage(aLog, ERROR_LOGGING_LEVEL, trim(where)//" : ", trim(msg))
end subroutine

!> logical(klog) -> integer(c_int): the bit pattern of .TRUE. is compiler dependent,
!> so the conversion happens on the Fortran side (once per solver New()).
pure function MaskToC(mask) result(imask)
	logical(klog), intent(in) :: mask(:)
	integer(c_int) :: imask(size(mask))
	imask = merge(1_c_int, 0_c_int, mask)
end function

end module
