! Fortran host side of the B200 GFN2-xTB engine: a TCalculator that forwards singlepoint through iso_c_binding to the
! C ABI of libxtb_b200.so (include/xtb_b200.h).  Written against the reference's calculator interface
! (src/type/calculator.f90:33-109, abstract interface `singlepoint`; src/xtb/calculator.f90:195-350 is the routine it
! replaces).  The build image has no Fortran compiler, so this file is shipped as source and is NOT compiled or tested
! here; the C ABI below it is what the test-suite exercises (tests/test_gpu_parity.py through ctypes).
!
! To use inside xtb: add this file to src/, select it in newCalculator (src/main/setup.f90:48-204) for
! `set%mode_extrun == p_ext_xtb` when a CUDA device is present, link with -lxtb_b200.
module xtb_b200_calculator
   use, intrinsic :: iso_c_binding
   use xtb_mctc_accuracy, only : wp
   use xtb_type_calculator, only : TCalculator
   use xtb_type_data, only : scc_results
   use xtb_type_environment, only : TEnvironment
   use xtb_type_molecule, only : TMolecule
   use xtb_type_restart, only : TRestart
   use xtb_solv_input, only : TSolvInput
   implicit none
   private

   public :: TB200Calculator, newB200Calculator

   !> GFN2-xTB calculator evaluated on a B200 through libxtb_b200
   type, extends(TCalculator) :: TB200Calculator
      type(c_ptr) :: batch = c_null_ptr
      real(wp) :: etemp = 300.0_wp
      integer :: maxiter = 250
      integer :: device = 0
   contains
      procedure :: singlepoint
      procedure :: writeInfo
      procedure :: setSolvent
      procedure :: releaseSolvent
      final :: delB200Calculator
   end type TB200Calculator

   interface
      function xtbb200_newBatch(nmol, natoms, numbers, charges, uhf, device) result(batch) &
            & bind(C, name="xtbb200_newBatch")
         import :: c_ptr, c_int, c_double
         integer(c_int), value :: nmol
         integer(c_int), intent(in) :: natoms(*), numbers(*)
         real(c_double), intent(in) :: charges(*)
         integer(c_int), intent(in) :: uhf(*)
         integer(c_int), value :: device
         type(c_ptr) :: batch
      end function xtbb200_newBatch
      subroutine xtbb200_delBatch(batch) bind(C, name="xtbb200_delBatch")
         import :: c_ptr
         type(c_ptr), intent(inout) :: batch
      end subroutine xtbb200_delBatch
      function xtbb200_setOptions(batch, accuracy, etemp, maxiter, guess) result(stat) &
            & bind(C, name="xtbb200_setOptions")
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: batch
         real(c_double), value :: accuracy, etemp
         integer(c_int), value :: maxiter, guess
         integer(c_int) :: stat
      end function xtbb200_setOptions
      function xtbb200_setWavefunction(batch, qsh, dipm, qp) result(stat) &
            & bind(C, name="xtbb200_setWavefunction")
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: batch
         real(c_double), intent(in) :: qsh(*), dipm(3, *), qp(6, *)
         integer(c_int) :: stat
      end function xtbb200_setWavefunction
      function xtbb200_singlepoint(batch, positions, restart, energies, gradients, charges, niter, status) &
            & result(stat) bind(C, name="xtbb200_singlepoint")
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: batch
         real(c_double), intent(in) :: positions(3, *)
         integer(c_int), value :: restart
         real(c_double), intent(out) :: energies(*), gradients(3, *), charges(*)
         integer(c_int), intent(out) :: niter(*), status(*)
         integer(c_int) :: stat
      end function xtbb200_singlepoint
      function xtbb200_getResult(batch, imol, out16) result(stat) bind(C, name="xtbb200_getResult")
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: batch
         integer(c_int), value :: imol
         real(c_double), intent(out) :: out16(16)
         integer(c_int) :: stat
      end function xtbb200_getResult
      function xtbb200_getArray(batch, name, out, len) result(stat) bind(C, name="xtbb200_getArray")
         import :: c_ptr, c_int, c_double, c_char, c_size_t
         type(c_ptr), value :: batch
         character(kind=c_char), intent(in) :: name(*)
         real(c_double), intent(out) :: out(*)
         integer(c_size_t), value :: len
         integer(c_int) :: stat
      end function xtbb200_getArray
      function xtbb200_setSolvent(batch, solvent, alpb, kernel, state, temperature, ion_strength, ion_rad) result(stat) &
            & bind(C, name="xtbb200_setSolvent")
         import :: c_ptr, c_int, c_double, c_char
         type(c_ptr), value :: batch
         character(kind=c_char), intent(in) :: solvent(*)
         integer(c_int), value :: alpb, kernel, state
         real(c_double), value :: temperature, ion_strength, ion_rad
         integer(c_int) :: stat
      end function xtbb200_setSolvent
      function xtbb200_releaseSolvent(batch, solvent, alpb, kernel, state, temperature, ion_strength, ion_rad) result(stat) &
            & bind(C, name="xtbb200_setSolvent")
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: batch, solvent   ! solvent = c_null_ptr releases the model
         integer(c_int), value :: alpb, kernel, state
         real(c_double), value :: temperature, ion_strength, ion_rad
         integer(c_int) :: stat
      end function xtbb200_releaseSolvent
      function xtbb200_lastError() result(msg) bind(C, name="xtbb200_lastError")
         import :: c_ptr
         type(c_ptr) :: msg
      end function xtbb200_lastError
   end interface

contains

!> Constructor: mirrors newXTBCalculator (src/xtb/calculator.f90:94-192) for method = 2
subroutine newB200Calculator(env, mol, calc, accuracy, device)
   character(len=*), parameter :: source = 'xtb_b200_newCalculator'
   type(TEnvironment), intent(inout) :: env
   type(TMolecule), intent(in) :: mol
   type(TB200Calculator), intent(out) :: calc
   real(wp), intent(in), optional :: accuracy
   integer, intent(in), optional :: device
   integer(c_int) :: natoms(1), uhf(1)
   real(c_double) :: charges(1)

   calc%accuracy = 1.0_wp
   if (present(accuracy)) calc%accuracy = accuracy
   if (present(device)) calc%device = device
   natoms(1) = mol%n
   charges(1) = mol%chrg
   uhf(1) = mol%uhf
   calc%batch = xtbb200_newBatch(1_c_int, natoms, int(mol%at, c_int), charges, uhf, int(calc%device, c_int))
   if (.not.c_associated(calc%batch)) then
      call env%error("Could not create the B200 engine (no CUDA device or unsupported element)", source)
   end if
end subroutine newB200Calculator

subroutine delB200Calculator(self)
   type(TB200Calculator), intent(inout) :: self
   if (c_associated(self%batch)) call xtbb200_delBatch(self%batch)
end subroutine delB200Calculator

!> Same contract as TxTBCalculator%singlepoint (src/xtb/calculator.f90:195-350): energy in Eh, gradient(3, n) in
!> Eh/Bohr, sigma zeroed (scf has no sigma argument), chk%wfn holds the converged q/qsh/dipm/qp afterwards.
subroutine singlepoint(self, env, mol, chk, printlevel, restart, energy, gradient, sigma, hlgap, results)
   character(len=*), parameter :: source = 'xtb_b200_singlepoint'
   class(TB200Calculator), intent(inout) :: self
   type(TEnvironment), intent(inout) :: env
   type(TMolecule), intent(inout) :: mol
   type(TRestart), intent(inout) :: chk
   integer, intent(in) :: printlevel
   logical, intent(in) :: restart
   real(wp), intent(out) :: energy
   real(wp), intent(out) :: gradient(:, :)
   real(wp), intent(out) :: sigma(:, :)
   real(wp), intent(out) :: hlgap
   type(scc_results), intent(out) :: results
   integer(c_int) :: stat, niter(1), status(1), irestart
   real(c_double) :: e(1), out16(16)
   real(c_double), allocatable :: g(:, :), q(:)

   call mol%update
   energy = 0.0_wp; gradient(:, :) = 0.0_wp; sigma(:, :) = 0.0_wp; hlgap = 0.0_wp
   stat = xtbb200_setOptions(self%batch, self%accuracy, self%etemp, int(self%maxiter, c_int), 0_c_int)
   irestart = 0
   if (restart .and. allocated(chk%wfn%qsh)) then
      stat = xtbb200_setWavefunction(self%batch, chk%wfn%qsh, chk%wfn%dipm, chk%wfn%qp)
      irestart = 1
   end if
   allocate(g(3, mol%n), q(mol%n))
   stat = xtbb200_singlepoint(self%batch, mol%xyz, irestart, e, g, q, niter, status)
   if (stat < 0) then
      call env%error("B200 engine failed", source)
      return
   end if
   energy = e(1)
   gradient(:, :) = g
   stat = xtbb200_getResult(self%batch, 0_c_int, out16)
   hlgap = out16(8)
   results%e_total = out16(1); results%e_elec = out16(2); results%e_rep = out16(3); results%e_es = out16(4)
   results%e_aes = out16(5); results%e_axc = out16(6); results%e_disp = out16(7); results%hl_gap = out16(8)
   results%gnorm = out16(9); results%dipole = out16(11:13)
   results%converged = status(1) == 0
   chk%wfn%q = q
   stat = xtbb200_getArray(self%batch, "qsh"//c_null_char, chk%wfn%qsh, int(size(chk%wfn%qsh), c_size_t))
   stat = xtbb200_getArray(self%batch, "dipm"//c_null_char, chk%wfn%dipm, int(size(chk%wfn%dipm), c_size_t))
   stat = xtbb200_getArray(self%batch, "qp"//c_null_char, chk%wfn%qp, int(size(chk%wfn%qp), c_size_t))
   if (status(1) /= 0) call env%error("Self consistent charge iterator did not converge", "scf")
end subroutine singlepoint

!> Implicit solvation: what addSolvationModel (src/main/setup.f90) does for TxTBCalculator with a TSolvInput
!> (src/solv/input.F90:30-66); a maintainer adds a `type is (TB200Calculator)` branch there that calls this routine.
!> input%kernel is the reference's gbKernel enumerator (1 Still, 2 P16), which is what the C ABI takes.
subroutine setSolvent(self, env, input)
   character(len=*), parameter :: source = 'xtb_b200_setSolvent'
   class(TB200Calculator), intent(inout) :: self
   type(TEnvironment), intent(inout) :: env
   type(TSolvInput), intent(in) :: input
   integer(c_int) :: stat
   if (input%cosmo .or. input%tmcosmo) then
      call env%error("COSMO is not available in the B200 GFN2-xTB engine", source)
      return
   end if
   stat = xtbb200_setSolvent(self%batch, input%solvent//c_null_char, merge(1_c_int, 0_c_int, input%alpb), &
      & int(input%kernel, c_int), int(input%state, c_int), real(input%temperature, c_double), &
      & real(input%ionStrength, c_double), real(input%ionRad, c_double))
   if (stat < 0) then
      call env%error("solvent: '"//input%solvent//"' is not parametrized", source)
      return
   end if
   self%lSolv = .true.
end subroutine setSolvent

subroutine releaseSolvent(self)
   class(TB200Calculator), intent(inout) :: self
   integer(c_int) :: stat
   stat = xtbb200_releaseSolvent(self%batch, c_null_ptr, 0_c_int, 2_c_int, 1_c_int, 298.15_c_double, 0.0_c_double, 0.0_c_double)
   self%lSolv = .false.
end subroutine releaseSolvent

subroutine writeInfo(self, unit, mol)
   class(TB200Calculator), intent(in) :: self
   integer, intent(in) :: unit
   type(TMolecule), intent(in) :: mol
   write(unit, '(a, i0, a)') "GFN2-xTB on B200 (libxtb_b200), ", mol%n, " atoms"
end subroutine writeInfo

end module xtb_b200_calculator
