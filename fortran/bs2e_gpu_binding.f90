! bs2e_gpu_binding.f90 -- ISO_C_BINDING layer between the reference's Fortran
! drivers and libbs2e_gpu.so (include/bs2e.h).
!
! Two modules:
!   bs2e_gpu_c  : bind(C) interfaces, one per C entry point that the Fortran
!                 side needs (the same style as the GSL interfaces in
!                 src/tools/wigner_tools.f90:7-26).
!   bs2e_gpu    : three subroutines with the NAMES and DUMMY-ARGUMENT LISTS of
!                 the reference routines they stand in for, so that
!                 src/apps/main_basis_setup.f90 keeps its three call lines
!                 (:80, :85, :108) unchanged:
!                   setup_Slater_integrals  (src/mat_els/mat_els.f90:172-178)
!                   compute_R_k_map         (src/tools/sparse_array_tools.f90:452-458)
!                   construct_block_tensor  (src/mat_els/hamiltonian.f90:106-114)
!
! The reference is compiled with -fdefault-integer-8 (CMakeLists.txt:44-56), so
! default integer = integer(c_int64_t) and default logical is 8 bytes; every
! integer crossing the boundary is converted explicitly anyway.
!
! NOTE: this image has no Fortran compiler (SURVEY.md F2); the file is written
! against the reference's type definitions and is exercised through the C ABI
! by the ctypes harness in b-spline-two-e_b200/bs2e (same call order, same
! array layouts).  See INTEGRATION.md for the build and use-line changes.
module bs2e_gpu_c
    use, intrinsic :: iso_c_binding
    implicit none

    interface
        function bs2e_last_error() bind(C, name="bs2e_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function bs2e_last_error

        function bs2e_ctx_create(k_spline, n_knots, knots, max_k, k_GL, gl_x, gl_w, device, ctx) &
                bind(C, name="bs2e_ctx_create") result(rc)
            import :: c_int, c_int64_t, c_double, c_ptr
            integer(c_int64_t), value :: k_spline, n_knots, max_k, k_GL, device
            real(c_double), intent(in) :: knots(*), gl_x(*), gl_w(*)
            type(c_ptr), intent(out) :: ctx
            integer(c_int) :: rc
        end function bs2e_ctx_create

        function bs2e_ctx_destroy(ctx) bind(C, name="bs2e_ctx_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: rc
        end function bs2e_ctx_destroy

        function bs2e_slater_cells(ctx) bind(C, name="bs2e_slater_cells") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: rc
        end function bs2e_slater_cells

        function bs2e_get_r_k(ctx, r_k, r_m_k, iv, i, j) bind(C, name="bs2e_get_r_k") result(rc)
            import :: c_int, c_int64_t, c_double, c_ptr
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: r_k(*), r_m_k(*)
            integer(c_int64_t), intent(out) :: iv(*), i(*), j(*)
            integer(c_int) :: rc
        end function bs2e_get_r_k

        function bs2e_get_r_d_k(ctx, r_d_k, iv, i, j, i_p, j_p) bind(C, name="bs2e_get_r_d_k") result(rc)
            import :: c_int, c_int64_t, c_double, c_ptr
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: r_d_k(*)
            integer(c_int64_t), intent(out) :: iv(*), i(*), j(*), i_p(*), j_p(*)
            integer(c_int) :: rc
        end function bs2e_get_r_d_k

        function bs2e_rk_build(ctx) bind(C, name="bs2e_rk_build") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: rc
        end function bs2e_rk_build

        function bs2e_rk_get(ctx, n_keys, keys, vals) bind(C, name="bs2e_rk_get") result(rc)
            import :: c_int, c_int64_t, c_double, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: n_keys
            integer(c_int64_t), intent(in) :: keys(*)
            real(c_double), intent(out) :: vals(*)
            integer(c_int) :: rc
        end function bs2e_rk_get

        function bs2e_set_one_particle(ctx, max_l_1p, H_vec, S) &
                bind(C, name="bs2e_set_one_particle") result(rc)
            import :: c_int, c_int64_t, c_double_complex, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: max_l_1p
            complex(c_double_complex), intent(in) :: H_vec(*), S(*)
            integer(c_int) :: rc
        end function bs2e_set_one_particle

        function bs2e_block_count(ctx, L, n_config, conf_n, conf_l, full, nnz_H, nnz_S) &
                bind(C, name="bs2e_block_count") result(rc)
            import :: c_int, c_int64_t, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: L, n_config, full
            integer(c_int64_t), intent(in) :: conf_n(2,*), conf_l(2,*)
            integer(c_int64_t), intent(out) :: nnz_H, nnz_S
            integer(c_int) :: rc
        end function bs2e_block_count

        function bs2e_block_fill(ctx, L, n_config, conf_n, conf_l, full, &
                                 H_ptr, H_idx, H_dat, S_ptr, S_idx, S_dat) &
                bind(C, name="bs2e_block_fill") result(rc)
            import :: c_int, c_int64_t, c_double_complex, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: L, n_config, full
            integer(c_int64_t), intent(in) :: conf_n(2,*), conf_l(2,*)
            integer(c_int64_t), intent(out) :: H_ptr(*), H_idx(*), S_ptr(*), S_idx(*)
            complex(c_double_complex), intent(out) :: H_dat(*), S_dat(*)
            integer(c_int) :: rc
        end function bs2e_block_fill
        function bs2e_set_radial_dipole(ctx, gauge, A, B) bind(C, name="bs2e_set_radial_dipole") result(rc)
            import :: c_ptr, c_int, c_int64_t, c_double_complex
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: gauge
            complex(c_double_complex), intent(in) :: A(*)
            type(c_ptr), value :: B                       ! r_inv_mat or c_null_ptr (length gauge)
            integer(c_int) :: rc
        end function bs2e_set_radial_dipole

        function bs2e_dip_block_count(ctx, q, sym1, n_config1, conf_n1, conf_l1, sym2, n_config2, conf_n2, &
                                      conf_l2, compute, nnz) bind(C, name="bs2e_dip_block_count") result(rc)
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: q, n_config1, n_config2, compute
            integer(c_int64_t), intent(in) :: sym1(3), sym2(3), conf_n1(*), conf_l1(*), conf_n2(*), conf_l2(*)
            integer(c_int64_t), intent(out) :: nnz
            integer(c_int) :: rc
        end function bs2e_dip_block_count

        function bs2e_dip_block_fill(ctx, q, sym1, n_config1, conf_n1, conf_l1, sym2, n_config2, conf_n2, &
                                     conf_l2, compute, index_ptr, indices, data) &
                                     bind(C, name="bs2e_dip_block_fill") result(rc)
            import :: c_ptr, c_int, c_int64_t, c_double_complex
            type(c_ptr), value :: ctx
            integer(c_int64_t), value :: q, n_config1, n_config2, compute
            integer(c_int64_t), intent(in) :: sym1(3), sym2(3), conf_n1(*), conf_l1(*), conf_n2(*), conf_l2(*)
            integer(c_int64_t), intent(out) :: index_ptr(*), indices(*)
            complex(c_double_complex), intent(out) :: data(*)
            integer(c_int) :: rc
        end function bs2e_dip_block_fill
    end interface

contains

    ! reference convention for failures is print + error stop
    ! (e.g. src/tools/sparse_array_tools.f90:527,553)
    subroutine bs2e_check(rc, where)
        use, intrinsic :: iso_fortran_env, only: stderr => error_unit
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        character(kind=c_char), pointer :: cmsg(:)
        character(len=512) :: msg
        integer :: q
        if (rc == 0) return
        call c_f_pointer(bs2e_last_error(), cmsg, [512])
        msg = ''
        do q = 1, 512
            if (cmsg(q) == c_null_char) exit
            msg(q:q) = cmsg(q)
        end do
        write(stderr,*) "bs2e_gpu, ", where, ": ", trim(msg)
        error stop
    end subroutine bs2e_check

end module bs2e_gpu_c


module bs2e_gpu
    use, intrinsic :: iso_c_binding
    use bs2e_gpu_c
    use bspline_tools, only: b_spline
    use quad_tools, only: setup_GL
    use sparse_array_tools, only: sparse_4d, sparse_6d, Nd_DOK, CSR_matrix
    use block_tools, only: block
    use orbital_tools, only: sym
    use omp_lib, only: omp_get_wtime
    implicit none
    private
    public :: setup_Slater_integrals, compute_R_k_map, construct_block_tensor, bs2e_gpu_finalize
    public :: construct_dip_block_tensor

    ! device-resident basis + R^k tensor; the Nd_DOK object of the caller stays
    ! an empty shell (nothing outside these three routines reads it,
    ! src/apps/main_basis_setup.f90:120)
    type(c_ptr), save :: ctx = c_null_ptr
    logical, save :: one_particle_set = .false.
    logical, save :: radial_dip_set = .false.

contains

    ! stands in for mat_els::setup_Slater_integrals (mat_els.f90:172-182).
    ! r_k, r_m_k, r_d_k arrive initialised (main_basis_setup.f90:77-79) and are
    ! filled in the reference's entry order so that any later Fortran reader
    ! finds what it expects.
    subroutine setup_Slater_integrals(b_splines, max_k, k_GL, r_k, r_m_k, r_d_k)
        type(b_spline), intent(in) :: b_splines
        integer, intent(in) :: max_k
        integer, intent(in) :: k_GL
        type(sparse_4d), intent(inout) :: r_k
        type(sparse_4d), intent(inout) :: r_m_k
        type(sparse_6d), intent(inout) :: r_d_k

        real(c_double), allocatable :: x(:), w(:)
        integer(c_int64_t) :: device
        character(len=32) :: env
        integer :: stat

        allocate(x(k_GL), w(k_GL))
        call setup_GL(k_GL, -1.d0, 1.d0, x, w)          ! the rule the reference integrates with
        device = 0
        call get_environment_variable("BS2E_DEVICE", env, status=stat)
        if (stat == 0) read(env,*) device
        if (c_associated(ctx)) call bs2e_check(bs2e_ctx_destroy(ctx), "ctx_destroy")
        call bs2e_check(bs2e_ctx_create(int(b_splines%k, c_int64_t), &
                                        int(size(b_splines%knots), c_int64_t), b_splines%knots, &
                                        int(max_k, c_int64_t), int(k_GL, c_int64_t), x, w, &
                                        device, ctx), "ctx_create")
        one_particle_set = .false.
        radial_dip_set = .false.   ! a new context holds no radial dipole matrices either
        call bs2e_check(bs2e_slater_cells(ctx), "slater_cells")
        call bs2e_check(bs2e_get_r_k(ctx, r_k%data, r_m_k%data, r_k%iv, r_k%i, r_k%j), "get_r_k")
        r_m_k%iv = r_k%iv
        r_m_k%i = r_k%i
        r_m_k%j = r_k%j
        call bs2e_check(bs2e_get_r_d_k(ctx, r_d_k%data, r_d_k%iv, r_d_k%i, r_d_k%j, &
                                       r_d_k%i_p, r_d_k%j_p), "get_r_d_k")
    end subroutine setup_Slater_integrals

    ! stands in for sparse_array_tools::compute_R_K_map (:452-493); the tensor
    ! stays on the device, R is left unallocated
    subroutine compute_R_k_map(r_d_k, r_k, r_m_k, b_splines, max_k, R)
        type(sparse_6d), intent(in) :: r_d_k
        type(sparse_4d), intent(in) :: r_k
        type(sparse_4d), intent(in) :: r_m_k
        type(b_spline), intent(in) :: b_splines
        integer, intent(in) :: max_k
        type(Nd_DOK), intent(inout) :: R
        double precision :: t_1, t_2

        t_1 = omp_get_wtime()
        call bs2e_check(bs2e_rk_build(ctx), "rk_build")
        t_2 = omp_get_wtime()
        R%N = 4
        R%N_val = max_k + 1
        write(6,*) "Time for construction (s): ", t_2 - t_1
    end subroutine compute_R_k_map

    ! stands in for hamiltonian::construct_block_tensor (:106-283): count_nnz,
    ! H_sp%init / S_sp%init on the Fortran side (allocatable components cannot
    ! be allocated from C), then the library fills index_ptr, indices, data.
    ! Called from inside "!$omp parallel do" (main_basis_setup.f90:105): the GPU
    ! serialises the blocks anyway, so the body is one critical section.
    subroutine construct_block_tensor(H, S, b_splines, term, max_k, R_k, H_sp, S_sp, full)
        type(block), dimension(:), allocatable, intent(in) :: H
        double complex, dimension(:,:), intent(in) :: S
        type(b_spline), intent(in) :: b_splines
        type(sym), intent(in) :: term
        integer, intent(in) :: max_k
        type(Nd_DOK), intent(inout) :: R_k
        type(CSR_matrix), intent(out) :: H_sp, S_sp
        logical, intent(in) :: full

        integer(c_int64_t), allocatable :: conf_n(:,:), conf_l(:,:)
        complex(c_double_complex), allocatable :: Hbuf(:,:,:), Sbuf(:,:)
        integer(c_int64_t) :: nnz_H, nnz_S, c_full
        integer :: i, l, n_b
        double precision :: t_1, t_2

        n_b = b_splines%n_b
        allocate(conf_n(2,term%n_config), conf_l(2,term%n_config))
        do i = 1, term%n_config
            conf_n(:,i) = int(term%configs(i)%n, c_int64_t)
            conf_l(:,i) = int(term%configs(i)%l, c_int64_t)
        end do
        c_full = merge(1_c_int64_t, 0_c_int64_t, full)

        !$omp critical(bs2e_gpu_block)
        t_1 = omp_get_wtime()
        if (.not. one_particle_set) then
            allocate(Hbuf(n_b,n_b,lbound(H,1):ubound(H,1)), Sbuf(n_b,n_b))
            do l = lbound(H,1), ubound(H,1)
                Hbuf(:,:,l) = H(l)%data
            end do
            Sbuf = S
            call bs2e_check(bs2e_set_one_particle(ctx, int(ubound(H,1)-lbound(H,1), c_int64_t), &
                                                  Hbuf, Sbuf), "set_one_particle")
            one_particle_set = .true.
        end if
        call bs2e_check(bs2e_block_count(ctx, int(term%l, c_int64_t), int(term%n_config, c_int64_t), &
                                         conf_n, conf_l, c_full, nnz_H, nnz_S), "block_count")
        call H_sp%init([term%n_config, term%n_config], int(nnz_H))
        call S_sp%init([term%n_config, term%n_config], int(nnz_S))
        write(6,*) term%l, term%m, term%pi, term%n_config
        call bs2e_check(bs2e_block_fill(ctx, int(term%l, c_int64_t), int(term%n_config, c_int64_t), &
                                        conf_n, conf_l, c_full, &
                                        H_sp%index_ptr, H_sp%indices, H_sp%data, &
                                        S_sp%index_ptr, S_sp%indices, S_sp%data), "block_fill")
        t_2 = omp_get_wtime()
        write(6,*) "Time to construct H_block (s): ", t_2 - t_1
        !$omp end critical(bs2e_gpu_block)
    end subroutine construct_block_tensor

    ! stands in for dipole::construct_dip_block_tensor (src/mat_els/dipole.f90:8-47), same
    ! dummies (type(radial_dipole) comes from mat_els).  Called from inside "!$omp parallel
    ! do" (main_basis_setup.f90:131-148): one critical section.  The overlap matrix is the
    ! one construct_block_tensor already sent (one_particle_set).
    subroutine construct_dip_block_tensor(syms, q, b_splines, S, radial_dip, dip_block, compute)
        use mat_els, only: radial_dipole
        type(sym), dimension(2), intent(in) :: syms
        integer, intent(in) :: q
        type(b_spline), intent(in) :: b_splines
        double complex, dimension(:,:), intent(in) :: S
        type(radial_dipole), intent(in) :: radial_dip
        type(CSR_matrix), intent(out) :: dip_block
        logical, intent(in) :: compute

        integer(c_int64_t), allocatable :: cn1(:,:), cl1(:,:), cn2(:,:), cl2(:,:)
        complex(c_double_complex), allocatable, target :: A(:,:), B(:,:)
        integer(c_int64_t) :: s1(3), s2(3), nnz, c_compute
        integer :: i

        allocate(cn1(2,syms(1)%n_config), cl1(2,syms(1)%n_config), cn2(2,syms(2)%n_config), cl2(2,syms(2)%n_config))
        do i = 1, syms(1)%n_config
            cn1(:,i) = int(syms(1)%configs(i)%n, c_int64_t)
            cl1(:,i) = int(syms(1)%configs(i)%l, c_int64_t)
        end do
        do i = 1, syms(2)%n_config
            cn2(:,i) = int(syms(2)%configs(i)%n, c_int64_t)
            cl2(:,i) = int(syms(2)%configs(i)%l, c_int64_t)
        end do
        s1 = [int(syms(1)%l, c_int64_t), int(syms(1)%m, c_int64_t), merge(1_c_int64_t, 0_c_int64_t, syms(1)%pi)]
        s2 = [int(syms(2)%l, c_int64_t), int(syms(2)%m, c_int64_t), merge(1_c_int64_t, 0_c_int64_t, syms(2)%pi)]
        c_compute = merge(1_c_int64_t, 0_c_int64_t, compute)

        !$omp critical(bs2e_gpu_block)
        if (.not. radial_dip_set) then
            if (radial_dip%gauge == 'l') then
                A = radial_dip%r_mat
                call bs2e_check(bs2e_set_radial_dipole(ctx, int(iachar('l'), c_int64_t), A, c_null_ptr), "set_radial_dipole")
            else
                A = radial_dip%dr_mat
                B = radial_dip%r_inv_mat
                call bs2e_check(bs2e_set_radial_dipole(ctx, int(iachar('v'), c_int64_t), A, c_loc(B)), "set_radial_dipole")
            end if
            radial_dip_set = .true.
        end if
        call bs2e_check(bs2e_dip_block_count(ctx, int(q, c_int64_t), s1, int(syms(1)%n_config, c_int64_t), cn1, cl1, &
                                             s2, int(syms(2)%n_config, c_int64_t), cn2, cl2, c_compute, nnz), &
                        "dip_block_count")
        call dip_block%init([syms(1)%n_config, syms(2)%n_config], int(nnz))
        if (nnz > 0) then
            call bs2e_check(bs2e_dip_block_fill(ctx, int(q, c_int64_t), s1, int(syms(1)%n_config, c_int64_t), cn1, cl1, &
                                                s2, int(syms(2)%n_config, c_int64_t), cn2, cl2, c_compute, &
                                                dip_block%index_ptr, dip_block%indices, dip_block%data), &
                            "dip_block_fill")
        end if
        !$omp end critical(bs2e_gpu_block)
    end subroutine construct_dip_block_tensor

    subroutine bs2e_gpu_finalize()
        if (c_associated(ctx)) call bs2e_check(bs2e_ctx_destroy(ctx), "ctx_destroy")
        ctx = c_null_ptr
        one_particle_set = .false.
        radial_dip_set = .false.
    end subroutine bs2e_gpu_finalize

end module bs2e_gpu
