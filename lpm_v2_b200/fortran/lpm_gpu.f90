!> @file lpm_gpu.f90
!> ISO_C_BINDING interface to liblpmgpu.so (include/lpm_gpu.h): the reference-side
!> binding a maintainer adds to lpm-v2's src/ (and to the source list in
!> src/CMakeLists.txt:1-16) to route the O(N^2) direct sums to the B200 library.
!> NOT compiled in this repository's image (no Fortran compiler is installed);
!> see INTEGRATION.md for the patched bodies of the private kernels that call it.
module LpmGpuModule
use iso_c_binding
use NumberKindsModule
use LoggerModule
implicit none
private
public :: LpmGpuInit, LpmGpuFinalize, LpmGpuCheck, MaskToC
public :: lpm_bve_velocity, lpm_bve_stream, lpm_plane_velocity, lpm_plane_stream
public :: lpm_betaplane_velocity, lpm_betaplane_stream
public :: lpm_pse_laplacian_sphere, lpm_pse_laplacian_plane
public :: lpm_pse_interpolate_sphere, lpm_pse_gradient_sphere, lpm_pse_divergence_sphere
public :: lpm_pse_gradient_plane, lpm_pse_second_partials_plane, lpm_pse_double_dot_plane
public :: lpm_swe_plane_rhs_integrals, lpm_swe_plane_velocity, lpm_swe_sphere_rhs_integrals
public :: lpm_gpu_pin, lpm_gpu_unpin
public :: lpm_gpu_init_rank, lpm_comm_unique_id, lpm_comm_init_rank, lpm_comm_alloc_shared, lpm_comm_free_shared
public :: lpm_bve_velocity_dev
public :: lpm_bve_solver_new, lpm_bve_solver_timestep, lpm_bve_solver_get_state, lpm_bve_solver_delete

interface
	integer(c_int) function lpm_gpu_init(ndev_requested, ndev_used) bind(C, name="lpm_gpu_init")
		import :: c_int
		integer(c_int), value :: ndev_requested
		integer(c_int), intent(out) :: ndev_used
	end function

	integer(c_int) function lpm_gpu_finalize() bind(C, name="lpm_gpu_finalize")
		import :: c_int
	end function

	type(c_ptr) function lpm_gpu_last_error() bind(C, name="lpm_gpu_last_error")
		import :: c_ptr
	end function

	integer(c_int) function lpm_gpu_pin(ptr, bytes) bind(C, name="lpm_gpu_pin")
		import :: c_int, c_ptr, c_int64_t
		type(c_ptr), value :: ptr
		integer(c_int64_t), value :: bytes
	end function

	integer(c_int) function lpm_gpu_unpin(ptr) bind(C, name="lpm_gpu_unpin")
		import :: c_int, c_ptr
		type(c_ptr), value :: ptr
	end function

	!> one MPI rank per GPU: claim one device, then join the communicator (id from rank 0 via MPI_BCAST)
	integer(c_int) function lpm_gpu_init_rank(device) bind(C, name="lpm_gpu_init_rank")
		import :: c_int
		integer(c_int), value :: device
	end function

	integer(c_int) function lpm_comm_unique_id(id) bind(C, name="lpm_comm_unique_id")
		import :: c_int, c_char
		character(kind=c_char), intent(out) :: id(128)
	end function

	integer(c_int) function lpm_comm_init_rank(world_size, rank, id) bind(C, name="lpm_comm_init_rank")
		import :: c_int, c_char
		integer(c_int), value :: world_size, rank
		character(kind=c_char), intent(in) :: id(128)
	end function

	!> collective: device memory every rank can store into over NVLink; a _dev sum whose outputs
	!> lie in such buffers delivers every slice to every rank (replaces the MPI_BCAST loop)
	integer(c_int) function lpm_comm_alloc_shared(bytes, ptr) bind(C, name="lpm_comm_alloc_shared")
		import :: c_int, c_int64_t, c_ptr
		integer(c_int64_t), value :: bytes
		type(c_ptr), intent(out) :: ptr
	end function

	integer(c_int) function lpm_comm_free_shared(ptr) bind(C, name="lpm_comm_free_shared")
		import :: c_int, c_ptr
		type(c_ptr), value :: ptr
	end function

	!> BVESphereVelocity on device-resident arrays: targets [ibeg, iend) (0-based, half open) =
	!> [indexStart(procRank) - 1, indexEnd(procRank)) of src/MPISetup.f90:138-144
	integer(c_int) function lpm_bve_velocity_dev(n, x, y, z, relvort, area, mask, radius, ibeg, iend, u, v, w, stream) &
			bind(C, name="lpm_bve_velocity_dev")
		import :: c_int, c_int64_t, c_double, c_ptr
		integer(c_int64_t), value :: n, ibeg, iend
		type(c_ptr), value :: x, y, z, relvort, area, mask, u, v, w, stream
		real(c_double), value :: radius
	end function

	!> replaces the loop nest + MPI_BCAST loop of BVESphereVelocity (SphereBVESolver.f90:395-429)
	integer(c_int) function lpm_bve_velocity(n, x, y, z, relvort, area, mask, radius, u, v, w) &
			bind(C, name="lpm_bve_velocity")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), relvort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: radius
		real(c_double), intent(out) :: u(*), v(*), w(*)
	end function

	!> SetStreamFunctionsOnMesh (SphereBVE.f90:445-485)
	integer(c_int) function lpm_bve_stream(n, x, y, z, relvort, absvort, area, mask, radius, relstream, absstream) &
			bind(C, name="lpm_bve_stream")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), relvort(*), absvort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: radius
		real(c_double), intent(out) :: relstream(*), absstream(*)
	end function

	!> planarIncompressibleVelocity (PlaneIncompressibleSolver.f90:278-316)
	integer(c_int) function lpm_plane_velocity(n, x, y, vort, area, mask, u, v) bind(C, name="lpm_plane_velocity")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), vort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), intent(out) :: u(*), v(*)
	end function

	!> SetStreamFunctionOnMesh (PlanarIncompressible.f90:470-505)
	integer(c_int) function lpm_plane_stream(n, x, y, vort, area, mask, psi) bind(C, name="lpm_plane_stream")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), vort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), intent(out) :: psi(*)
	end function

	!> BetaPlaneVelocity (BetaPlaneSolver.f90:227-267)
	integer(c_int) function lpm_betaplane_velocity(n, x, y, relvort, area, mask, u, v) &
			bind(C, name="lpm_betaplane_velocity")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), relvort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), intent(out) :: u(*), v(*)
	end function

	!> SetStreamFunctionsOnMesh (BetaPlane.f90:399-442)
	integer(c_int) function lpm_betaplane_stream(n, x, y, relvort, absvort, area, mask, relstream, absstream) &
			bind(C, name="lpm_betaplane_stream")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), relvort(*), absvort(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), intent(out) :: relstream(*), absstream(*)
	end function

	!> PSESphereLaplacianAtParticles (PSEDirectSum.f90:502-535)
	integer(c_int) function lpm_pse_laplacian_sphere(n, x, y, z, f, area, mask, eps, sphere_radius, lap) &
			bind(C, name="lpm_pse_laplacian_sphere")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), f(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps, sphere_radius
		real(c_double), intent(out) :: lap(*)
	end function

	!> PSEPlaneLaplacianAtParticles (PSEDirectSum.f90:467-500)
	integer(c_int) function lpm_pse_laplacian_plane(n, x, y, f, area, mask, eps, lap) &
			bind(C, name="lpm_pse_laplacian_plane")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), f(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps
		real(c_double), intent(out) :: lap(*)
	end function

	!> PSESphereInterpolateScalar (PSEDirectSum.f90:151-168) at m locations
	integer(c_int) function lpm_pse_interpolate_sphere(n, x, y, z, f, area, mask, eps, sphere_radius, &
			m, tx, ty, tz, fOut) bind(C, name="lpm_pse_interpolate_sphere")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n, m
		real(c_double), intent(in) :: x(*), y(*), z(*), f(*), area(*), tx(*), ty(*), tz(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps, sphere_radius
		real(c_double), intent(out) :: fOut(*)
	end function

	!> PSESphereGradientAtParticles (PSEDirectSum.f90:221-267)
	integer(c_int) function lpm_pse_gradient_sphere(n, x, y, z, f, area, mask, eps, sphere_radius, gx, gy, gz) &
			bind(C, name="lpm_pse_gradient_sphere")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), f(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps, sphere_radius
		real(c_double), intent(out) :: gx(*), gy(*), gz(*)
	end function

	!> PSESphereDivergenceAtParticles (PSEDirectSum.f90:537-579)
	integer(c_int) function lpm_pse_divergence_sphere(n, x, y, z, u, v, w, area, mask, eps, sphere_radius, div) &
			bind(C, name="lpm_pse_divergence_sphere")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), u(*), v(*), w(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps, sphere_radius
		real(c_double), intent(out) :: div(*)
	end function

	!> PSEPlaneGradientAtParticles (PSEDirectSum.f90:180-218)
	integer(c_int) function lpm_pse_gradient_plane(n, x, y, f, area, mask, eps, gx, gy) &
			bind(C, name="lpm_pse_gradient_plane")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), f(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps
		real(c_double), intent(out) :: gx(*), gy(*)
	end function

	!> PSEPlaneSecondPartialsAtParticles (PSEDirectSum.f90:269-320)
	integer(c_int) function lpm_pse_second_partials_plane(n, x, y, gx, gy, area, mask, eps, dxx, dxy, dyy) &
			bind(C, name="lpm_pse_second_partials_plane")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), gx(*), gy(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps
		real(c_double), intent(out) :: dxx(*), dxy(*), dyy(*)
	end function

	!> PSEPlaneDoubleDotProductAtParticles (PSEDirectSum.f90:322-365)
	integer(c_int) function lpm_pse_double_dot_plane(n, x, y, u, v, area, mask, eps, dd) &
			bind(C, name="lpm_pse_double_dot_plane")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), u(*), v(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: eps
		real(c_double), intent(out) :: dd(*)
	end function

	!> SWEPlaneRHSIntegrals (SWEPlaneSolver.f90:457-560); surf = h + topoFn(x, y)
	integer(c_int) function lpm_swe_plane_rhs_integrals(n, x, y, vort, div, surf, area, mask, pse_eps, &
			u, v, doubleDot, lapSurf) bind(C, name="lpm_swe_plane_rhs_integrals")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), vort(*), div(*), surf(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: pse_eps
		real(c_double), intent(out) :: u(*), v(*), doubleDot(*), lapSurf(*)
	end function

	!> SetVelocityFromFieldData (PlanarSWE.f90:469-494)
	integer(c_int) function lpm_swe_plane_velocity(n, x, y, vort, div, area, mask, u, v) &
			bind(C, name="lpm_swe_plane_velocity")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), vort(*), div(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), intent(out) :: u(*), v(*)
	end function

	!> SWESphereRHSIntegrals (SphereSWESolver.f90:296-375); surf = h + topoFn(x, y, z)
	integer(c_int) function lpm_swe_sphere_rhs_integrals(n, x, y, z, vort, div, surf, area, mask, radius, pse_eps, &
			u, v, w, doubleDot, lapSurf) bind(C, name="lpm_swe_sphere_rhs_integrals")
		import :: c_int, c_int64_t, c_double
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), vort(*), div(*), surf(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: radius, pse_eps
		real(c_double), intent(out) :: u(*), v(*), w(*), doubleDot(*), lapSurf(*)
	end function

	!> device-resident BVESolver: New / Timestep / Delete (SphereBVESolver.f90:112-168, 219-353)
	integer(c_int) function lpm_bve_solver_new(n, x, y, z, relvort, absvort, u, v, w, area, mask, radius, &
			rotation_rate, handle) bind(C, name="lpm_bve_solver_new")
		import :: c_int, c_int64_t, c_double, c_ptr
		integer(c_int64_t), value :: n
		real(c_double), intent(in) :: x(*), y(*), z(*), relvort(*), absvort(*), u(*), v(*), w(*), area(*)
		integer(c_int), intent(in) :: mask(*)
		real(c_double), value :: radius, rotation_rate
		type(c_ptr), intent(out) :: handle
	end function

	integer(c_int) function lpm_bve_solver_timestep(handle, dt, with_stream) bind(C, name="lpm_bve_solver_timestep")
		import :: c_int, c_double, c_ptr
		type(c_ptr), value :: handle
		real(c_double), value :: dt
		integer(c_int), value :: with_stream
	end function

	integer(c_int) function lpm_bve_solver_get_state(handle, x, y, z, relvort, u, v, w, relstream, absstream) &
			bind(C, name="lpm_bve_solver_get_state")
		import :: c_int, c_double, c_ptr
		type(c_ptr), value :: handle
		real(c_double), intent(out) :: x(*), y(*), z(*), relvort(*), u(*), v(*), w(*), relstream(*), absstream(*)
	end function

	integer(c_int) function lpm_bve_solver_delete(handle) bind(C, name="lpm_bve_solver_delete")
		import :: c_int, c_ptr
		type(c_ptr), value :: handle
	end function
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
