! rrtmg_lw_b200_shim.f90 -- thin ISO_C_BINDING layer that keeps the reference's GCM interface
! UNCHANGED and routes it to the B200 library (include/rrtmg_lw_b200.h).
!
! The shim provides replacements for three reference modules, with the same module names, public names
! and dummy-argument lists, so a host GCM relinks without touching its radiation driver:
!
!   module rrtmg_lw_init        : rrtmg_lw_ini(cpdair)             src/rrtmg_lw_init.f90:47            (this file)
!   module mcica_subcol_gen_lw  : get_alpha, mcica_subcol_lw       src/mcica_subcol_gen_lw.f90:68,183  (this file)
!   module rrtmg_lw_rad         : rrtmg_lw(ncol,nlay,icld,...)     one of, like the reference's two source files:
!        rrtmg_lw_b200_shim_rad.f90          McICA sub-column arrays    src/rrtmg_lw_rad.f90:99-108
!        rrtmg_lw_b200_shim_rad_nomcica.f90  grid-mean cloud arrays     src/rrtmg_lw_rad.nomcica.f90:99-108
!
! Build: compile this file and ONE of the two rad files INSTEAD OF the reference's rrtmg_lw_init.f90 (its table builders are
! re-done inside the library), rrtmg_lw_rad.f90, mcica_subcol_gen_lw.f90, rrtmg_lw_setcoef.f90,
! rrtmg_lw_taumol.f90, rrtmg_lw_cldprmc.f90, rrtmg_lw_rtrnmc.f90, mcica_random_numbers.f90; keep
! modules/parkind.f90, modules/rrlw_kg01..16.f90 and the k-data source (rrtmg_lw_k_g.f90 or
! rrtmg_lw_read_nc.f90) -- the shim calls the reference's own lw_kgb01..16 and hands the filled
! 16-g arrays to the library.  Link with -lrrtmg_lw_b200 -lcudart.
!
! STATUS: written against the C header; NOT compiled in the development image (no Fortran
! compiler there, DESIGN.md "Environment limits").  The Python ctypes mirror
! (rrtmg_lw_b200/api.py) binds the very same entry points and is what the tests exercise.
!
! All arrays are passed by address without copies when they are contiguous and share one leading
! (column) extent; `ld` = size(x,1) so GCM chunks with pcols > ncol work.  A non-zero return code
! is turned into `error stop` with the reference's STOP text (rrtmg_lw_b200_strerror).

      module rrtmg_lw_b200_c
      use iso_c_binding
      implicit none
      integer(c_int), parameter :: RRTMG_LW_HOST = 0

      interface
         integer(c_int) function c_kdata_put(band, name, data, n) bind(C, name='rrtmg_lw_b200_kdata_put')
         import :: c_int, c_long, c_char, c_double
         integer(c_int), value :: band
         character(kind=c_char), intent(in) :: name(*)
         real(c_double), intent(in) :: data(*)
         integer(c_long), value :: n
         end function

         integer(c_int) function c_ini(cpdair, kdata, nkdata, default_real_kind) bind(C, name='rrtmg_lw_b200_ini')
         import :: c_int, c_long, c_double, c_ptr
         real(c_double), value :: cpdair
         type(c_ptr), value :: kdata
         integer(c_long), value :: nkdata
         integer(c_int), value :: default_real_kind
         end function

         integer(c_int) function c_rrtmg_lw(ncol, ld, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, &
               h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, &
               inflglw, iceflglw, liqflglw, cldfmcl, taucmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, tauaer, &
               uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, memspace, stream) &
               bind(C, name='rrtmg_lw_b200_rrtmg_lw')
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: ncol, ld, nlay, idrv, inflglw, iceflglw, liqflglw, memspace
         integer(c_int), intent(inout) :: icld
         real(c_double), intent(in) :: play(*), plev(*), tlay(*), tlev(*), tsfc(*), h2ovmr(*), o3vmr(*), &
               co2vmr(*), ch4vmr(*), n2ovmr(*), o2vmr(*), cfc11vmr(*), cfc12vmr(*), cfc22vmr(*), ccl4vmr(*), &
               emis(*), cldfmcl(*), taucmcl(*), ciwpmcl(*), clwpmcl(*), reicmcl(*), relqmcl(*), tauaer(*)
         real(c_double), intent(out) :: uflx(*), dflx(*), hr(*), uflxc(*), dflxc(*), hrc(*)
         type(c_ptr), value :: duflx_dt, duflxc_dt, stream
         end function

         integer(c_int) function c_rrtmg_lw_nomcica(ncol, ld, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, &
               h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, &
               inflglw, iceflglw, liqflglw, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer, &
               uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt, memspace, stream) &
               bind(C, name='rrtmg_lw_b200_rrtmg_lw_nomcica')
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: ncol, ld, nlay, idrv, inflglw, iceflglw, liqflglw, memspace
         integer(c_int), intent(inout) :: icld
         real(c_double), intent(in) :: play(*), plev(*), tlay(*), tlev(*), tsfc(*), h2ovmr(*), o3vmr(*), &
               co2vmr(*), ch4vmr(*), n2ovmr(*), o2vmr(*), cfc11vmr(*), cfc12vmr(*), cfc22vmr(*), ccl4vmr(*), &
               emis(*), cldfr(*), taucld(*), cicewp(*), cliqwp(*), reice(*), reliq(*), tauaer(*)
         real(c_double), intent(out) :: uflx(*), dflx(*), hr(*), uflxc(*), dflxc(*), hrc(*)
         type(c_ptr), value :: duflx_dt, duflxc_dt, stream
         end function

         integer(c_int) function c_mcica_subcol_lw(iplon, ncol, ld, nlay, icld, permuteseed, irng, play, &
               cldfrac, ciwp, clwp, rei, rel, tauc, alpha, cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, &
               taucmcl, memspace, stream) bind(C, name='rrtmg_lw_b200_mcica_subcol_lw')
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: iplon, ncol, ld, nlay, icld, permuteseed, memspace
         integer(c_int), intent(inout) :: irng
         real(c_double), intent(in) :: play(*), cldfrac(*), ciwp(*), clwp(*), rei(*), rel(*), tauc(*), alpha(*)
         real(c_double), intent(out) :: cldfmcl(*), ciwpmcl(*), clwpmcl(*), reicmcl(*), relqmcl(*), taucmcl(*)
         type(c_ptr), value :: stream
         end function

         integer(c_int) function c_get_alpha(iplon, ncol, ld, nlayers, icld, idcor, decorr_con, dz, lat, &
               juldat, cldfrac, alpha, memspace, stream) bind(C, name='rrtmg_lw_b200_get_alpha')
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: iplon, ncol, ld, nlayers, icld, idcor, juldat, memspace
         real(c_double), value :: decorr_con
         real(c_double), intent(in) :: dz(*), lat(*), cldfrac(*)
         real(c_double), intent(inout) :: alpha(*)
         type(c_ptr), value :: stream
         end function

         type(c_ptr) function c_strerror(code) bind(C, name='rrtmg_lw_b200_strerror')
         import :: c_int, c_ptr
         integer(c_int), value :: code
         end function

         ! the GPU this thread's calls use (one rank per GPU: call once after MPI_Init with the node-local rank)
         integer(c_int) function c_set_device(device) bind(C, name='rrtmg_lw_b200_set_device')
         import :: c_int
         integer(c_int), value :: device
         end function

         ! multi-GPU output gather (one rank per GPU; include/rrtmg_lw_b200.h): rank 0 makes the 128-byte id, the host
         ! broadcasts it (MPI_Bcast(id, 128, MPI_BYTE, 0, comm)), every rank calls comm_init, then gathers per step
         integer(c_int) function c_comm_unique_id(id128) bind(C, name='rrtmg_lw_b200_comm_unique_id')
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id128(128)
         end function
         integer(c_int) function c_comm_init(nranks, rank, id128) bind(C, name='rrtmg_lw_b200_comm_init')
         import :: c_int, c_char
         integer(c_int), value :: nranks, rank
         character(kind=c_char), intent(in) :: id128(128)
         end function
         integer(c_int) function c_gather(send, count, recv, root, stream) bind(C, name='rrtmg_lw_b200_gather')
         import :: c_int, c_long, c_ptr
         type(c_ptr), value :: send, recv, stream      ! DEVICE addresses (c_loc of device arrays / host_data use_device)
         integer(c_long), value :: count
         integer(c_int), value :: root
         end function
         integer(c_int) function c_comm_finalize() bind(C, name='rrtmg_lw_b200_comm_finalize')
         import :: c_int
         end function

         ! run-time switches without a reference counterpart (include/rrtmg_lw_b200.h, INTEGRATION.md section 6)
         integer(c_int) function c_set_option(name, value) bind(C, name='rrtmg_lw_b200_set_option')
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), value :: value
         end function
      end interface

      contains

      ! irng = 1 on a column-sharded batch (one MPI rank per GPU): this rank's calls carry columns
      ! [col_offset+1, col_offset+ncol] of a batch of global_ncol columns; the library steps the batch's single
      ! Mersenne-Twister stream (src/mcica_subcol_gen_lw.f90:523-530) over the other ranks' words.
      ! global_ncol = 0 switches it off.
      subroutine rrtmg_lw_b200_set_mt_block(global_ncol, col_offset)
      integer, intent(in) :: global_ncol, col_offset
      call check(c_set_option('mt_global_ncol'//c_null_char, int(global_ncol, c_int)), 'set_option(mt_global_ncol)')
      call check(c_set_option('mt_col_offset'//c_null_char, int(col_offset, c_int)), 'set_option(mt_col_offset)')
      end subroutine rrtmg_lw_b200_set_mt_block

      subroutine rrtmg_lw_b200_set_device(device)
      integer, intent(in) :: device
      call check(c_set_device(int(device, c_int)), 'set_device')
      end subroutine rrtmg_lw_b200_set_device

      ! The reference also lists inatm (module rrtmg_lw_rad) and generate_stochastic_clouds (module
      ! mcica_subcol_gen_lw) as public.  Both are per-column internals of its CPU pipeline -- inatm fills one column's
      ! work arrays for setcoef / taumol, generate_stochastic_clouds is the body of mcica_subcol_lw on Pa pressures --
      ! whose outputs never leave the device here.  The shim keeps the names (a `use ..., only:` list that mentions
      ! them still compiles) and stops with this message if a host model really calls one.
      subroutine internal_only(name)
      character(len=*), intent(in) :: name
      write(*,'(a,a,a)') ' ', name, ': internal to the B200 library (the column state stays on the device); '// &
            'call rrtmg_lw / mcica_subcol_lw instead'
      error stop 1
      end subroutine internal_only

      ! reference behaviour on these conditions is STOP 'text' (e.g. src/rrtmg_lw_cldprmc.f90:204)
      subroutine check(rc, where)
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      character(len=160) :: text
      integer :: i
      if (rc == 0) return
      call c_f_pointer(c_strerror(rc), msg, [160])
      text = ' '
      do i = 1, 160
         if (msg(i) == c_null_char) exit
         text(i:i) = msg(i)
      end do
      write(*,'(a,a,a,a)') ' ', where, ': ', trim(text)
      error stop 1
      end subroutine check

      end module rrtmg_lw_b200_c

!------------------------------------------------------------------------------------------------
      module rrtmg_lw_init
      use iso_c_binding
      use parkind, only : im => kind_im, rb => kind_rb
      use rrtmg_lw_b200_c
      implicit none
      public :: rrtmg_lw_ini

      contains

      subroutine rrtmg_lw_ini(cpdair)
      real(kind=rb), intent(in) :: cpdair
      external lw_kgb01, lw_kgb02, lw_kgb03, lw_kgb04, lw_kgb05, lw_kgb06, lw_kgb07, lw_kgb08, &
               lw_kgb09, lw_kgb10, lw_kgb11, lw_kgb12, lw_kgb13, lw_kgb14, lw_kgb15, lw_kgb16
      integer(c_int) :: kind_default
      real :: probe
      ! the reference's own data routines (src/rrtmg_lw_init.f90:99-114) fill modules rrlw_kgNN
      call lw_kgb01; call lw_kgb02; call lw_kgb03; call lw_kgb04
      call lw_kgb05; call lw_kgb06; call lw_kgb07; call lw_kgb08
      call lw_kgb09; call lw_kgb10; call lw_kgb11; call lw_kgb12
      call lw_kgb13; call lw_kgb14; call lw_kgb15; call lw_kgb16
      call put_band01; call put_band02; call put_band03; call put_band04
      call put_band05; call put_band06; call put_band07; call put_band08
      call put_band09; call put_band10; call put_band11; call put_band12
      call put_band13; call put_band14; call put_band15; call put_band16
      ! default-real kind of THIS build (4 plain, 8 under -r8 / -fdefault-real-8): selects the value of
      ! the reference's default-kind literals inside the library (DESIGN.md "Default real kind")
      kind_default = int(kind(probe), c_int)
      call check(c_ini(real(cpdair, c_double), c_null_ptr, 0_c_long, kind_default), 'RRTMG_LW_INI')
      end subroutine rrtmg_lw_ini

      subroutine put(band, name, a, n)
      integer, intent(in) :: band
      character(len=*), intent(in) :: name
      real(kind=rb), intent(in) :: a(*)
      integer(c_long), intent(in) :: n
      call check(c_kdata_put(int(band, c_int), trim(name)//c_null_char, a, n), 'RRTMG_LW_INI')
      end subroutine put

      subroutine put_band01
      use rrlw_kg01, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mn2, kbo_mn2
      call put(1, 'kao', kao, int(size(kao), c_long))
      call put(1, 'kbo', kbo, int(size(kbo), c_long))
      call put(1, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(1, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(1, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(1, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(1, 'kao_mn2', kao_mn2, int(size(kao_mn2), c_long))
      call put(1, 'kbo_mn2', kbo_mn2, int(size(kbo_mn2), c_long))
      end subroutine put_band01

      subroutine put_band02
      use rrlw_kg02, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo
      call put(2, 'kao', kao, int(size(kao), c_long))
      call put(2, 'kbo', kbo, int(size(kbo), c_long))
      call put(2, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(2, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(2, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(2, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      end subroutine put_band02

      subroutine put_band03
      use rrlw_kg03, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mn2o, kbo_mn2o
      call put(3, 'kao', kao, int(size(kao), c_long))
      call put(3, 'kbo', kbo, int(size(kbo), c_long))
      call put(3, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(3, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(3, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(3, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(3, 'kao_mn2o', kao_mn2o, int(size(kao_mn2o), c_long))
      call put(3, 'kbo_mn2o', kbo_mn2o, int(size(kbo_mn2o), c_long))
      end subroutine put_band03

      subroutine put_band04
      use rrlw_kg04, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo
      call put(4, 'kao', kao, int(size(kao), c_long))
      call put(4, 'kbo', kbo, int(size(kbo), c_long))
      call put(4, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(4, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(4, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(4, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      end subroutine put_band04

      subroutine put_band05
      use rrlw_kg05, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mo3, ccl4o
      call put(5, 'kao', kao, int(size(kao), c_long))
      call put(5, 'kbo', kbo, int(size(kbo), c_long))
      call put(5, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(5, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(5, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(5, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(5, 'kao_mo3', kao_mo3, int(size(kao_mo3), c_long))
      call put(5, 'ccl4o', ccl4o, int(size(ccl4o), c_long))
      end subroutine put_band05

      subroutine put_band06
      use rrlw_kg06, only : kao, selfrefo, forrefo, fracrefao, kao_mco2, cfc11adjo, cfc12o
      call put(6, 'kao', kao, int(size(kao), c_long))
      call put(6, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(6, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(6, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(6, 'kao_mco2', kao_mco2, int(size(kao_mco2), c_long))
      call put(6, 'cfc11adjo', cfc11adjo, int(size(cfc11adjo), c_long))
      call put(6, 'cfc12o', cfc12o, int(size(cfc12o), c_long))
      end subroutine put_band06

      subroutine put_band07
      use rrlw_kg07, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mco2, kbo_mco2
      call put(7, 'kao', kao, int(size(kao), c_long))
      call put(7, 'kbo', kbo, int(size(kbo), c_long))
      call put(7, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(7, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(7, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(7, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(7, 'kao_mco2', kao_mco2, int(size(kao_mco2), c_long))
      call put(7, 'kbo_mco2', kbo_mco2, int(size(kbo_mco2), c_long))
      end subroutine put_band07

      subroutine put_band08
      use rrlw_kg08, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mco2, kao_mn2o, kao_mo3, kbo_mco2, kbo_mn2o, cfc12o, cfc22adjo
      call put(8, 'kao', kao, int(size(kao), c_long))
      call put(8, 'kbo', kbo, int(size(kbo), c_long))
      call put(8, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(8, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(8, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(8, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(8, 'kao_mco2', kao_mco2, int(size(kao_mco2), c_long))
      call put(8, 'kao_mn2o', kao_mn2o, int(size(kao_mn2o), c_long))
      call put(8, 'kao_mo3', kao_mo3, int(size(kao_mo3), c_long))
      call put(8, 'kbo_mco2', kbo_mco2, int(size(kbo_mco2), c_long))
      call put(8, 'kbo_mn2o', kbo_mn2o, int(size(kbo_mn2o), c_long))
      call put(8, 'cfc12o', cfc12o, int(size(cfc12o), c_long))
      call put(8, 'cfc22adjo', cfc22adjo, int(size(cfc22adjo), c_long))
      end subroutine put_band08

      subroutine put_band09
      use rrlw_kg09, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mn2o, kbo_mn2o
      call put(9, 'kao', kao, int(size(kao), c_long))
      call put(9, 'kbo', kbo, int(size(kbo), c_long))
      call put(9, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(9, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(9, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(9, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(9, 'kao_mn2o', kao_mn2o, int(size(kao_mn2o), c_long))
      call put(9, 'kbo_mn2o', kbo_mn2o, int(size(kbo_mn2o), c_long))
      end subroutine put_band09

      subroutine put_band10
      use rrlw_kg10, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo
      call put(10, 'kao', kao, int(size(kao), c_long))
      call put(10, 'kbo', kbo, int(size(kbo), c_long))
      call put(10, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(10, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(10, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(10, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      end subroutine put_band10

      subroutine put_band11
      use rrlw_kg11, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo, kao_mo2, kbo_mo2
      call put(11, 'kao', kao, int(size(kao), c_long))
      call put(11, 'kbo', kbo, int(size(kbo), c_long))
      call put(11, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(11, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(11, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(11, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(11, 'kao_mo2', kao_mo2, int(size(kao_mo2), c_long))
      call put(11, 'kbo_mo2', kbo_mo2, int(size(kbo_mo2), c_long))
      end subroutine put_band11

      subroutine put_band12
      use rrlw_kg12, only : kao, selfrefo, forrefo, fracrefao
      call put(12, 'kao', kao, int(size(kao), c_long))
      call put(12, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(12, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(12, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      end subroutine put_band12

      subroutine put_band13
      use rrlw_kg13, only : kao, selfrefo, forrefo, fracrefao, fracrefbo, kao_mco2, kao_mco, kbo_mo3
      call put(13, 'kao', kao, int(size(kao), c_long))
      call put(13, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(13, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(13, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(13, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      call put(13, 'kao_mco2', kao_mco2, int(size(kao_mco2), c_long))
      call put(13, 'kao_mco', kao_mco, int(size(kao_mco), c_long))
      call put(13, 'kbo_mo3', kbo_mo3, int(size(kbo_mo3), c_long))
      end subroutine put_band13

      subroutine put_band14
      use rrlw_kg14, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo
      call put(14, 'kao', kao, int(size(kao), c_long))
      call put(14, 'kbo', kbo, int(size(kbo), c_long))
      call put(14, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(14, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(14, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(14, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      end subroutine put_band14

      subroutine put_band15
      use rrlw_kg15, only : kao, selfrefo, forrefo, fracrefao, kao_mn2
      call put(15, 'kao', kao, int(size(kao), c_long))
      call put(15, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(15, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(15, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(15, 'kao_mn2', kao_mn2, int(size(kao_mn2), c_long))
      end subroutine put_band15

      subroutine put_band16
      use rrlw_kg16, only : kao, kbo, selfrefo, forrefo, fracrefao, fracrefbo
      call put(16, 'kao', kao, int(size(kao), c_long))
      call put(16, 'kbo', kbo, int(size(kbo), c_long))
      call put(16, 'selfrefo', selfrefo, int(size(selfrefo), c_long))
      call put(16, 'forrefo', forrefo, int(size(forrefo), c_long))
      call put(16, 'fracrefao', fracrefao, int(size(fracrefao), c_long))
      call put(16, 'fracrefbo', fracrefbo, int(size(fracrefbo), c_long))
      end subroutine put_band16
      end module rrtmg_lw_init

!------------------------------------------------------------------------------------------------
      module mcica_subcol_gen_lw
      use iso_c_binding
      use parkind, only : im => kind_im, rb => kind_rb
      use rrtmg_lw_b200_c
      implicit none
      public :: get_alpha, mcica_subcol_lw, generate_stochastic_clouds

      contains

      ! src/mcica_subcol_gen_lw.f90:295-296 -- name kept, see rrtmg_lw_b200_c::internal_only
      subroutine generate_stochastic_clouds(ncol, nlay, nsubcol, icld, irng, pmid, cld, clwp, ciwp, &
                               alpha, tauc, cld_stoch, clwp_stoch, ciwp_stoch, tauc_stoch, changeSeed)
      integer(kind=im), intent(in) :: ncol, nlay, nsubcol, icld
      integer(kind=im), intent(inout) :: irng
      real(kind=rb), intent(in) :: pmid(:,:), cld(:,:), clwp(:,:), ciwp(:,:), alpha(:,:), tauc(:,:,:)
      real(kind=rb), intent(out) :: cld_stoch(:,:,:), clwp_stoch(:,:,:), ciwp_stoch(:,:,:), tauc_stoch(:,:,:)
      integer(kind=im), optional, intent(in) :: changeSeed
      call internal_only('GENERATE_STOCHASTIC_CLOUDS')
      end subroutine generate_stochastic_clouds

      subroutine get_alpha(iplon, ncol, nlayers, icld, idcor, decorr_con, &
                           dz, lat, juldat, cldfrac, alpha)
      integer(kind=im), intent(in) :: iplon, ncol, nlayers, icld, idcor, juldat
      real(kind=rb), intent(in) :: decorr_con
      real(kind=rb), intent(in), contiguous :: dz(:,:), lat(:), cldfrac(:,:)
      real(kind=rb), intent(out), contiguous :: alpha(:,:)
      call check(c_get_alpha(int(iplon,c_int), int(ncol,c_int), int(size(dz,1),c_int), int(nlayers,c_int), &
            int(icld,c_int), int(idcor,c_int), real(decorr_con,c_double), dz, lat, int(juldat,c_int), cldfrac, &
            alpha, RRTMG_LW_HOST, c_null_ptr), 'GET_ALPHA')
      end subroutine get_alpha

      subroutine mcica_subcol_lw(iplon, ncol, nlay, icld, permuteseed, irng, play, &
                       cldfrac, ciwp, clwp, rei, rel, tauc, alpha, cldfmcl, &
                       ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl)
      integer(kind=im), intent(in) :: iplon, ncol, nlay, icld, permuteseed
      integer(kind=im), intent(inout) :: irng
      real(kind=rb), intent(in), contiguous :: play(:,:), cldfrac(:,:), tauc(:,:,:), ciwp(:,:), clwp(:,:), &
                                               rei(:,:), rel(:,:), alpha(:,:)
      real(kind=rb), intent(out), contiguous :: cldfmcl(:,:,:), ciwpmcl(:,:,:), clwpmcl(:,:,:), relqmcl(:,:), &
                                                reicmcl(:,:), taucmcl(:,:,:)
      integer(c_int) :: ir
      ir = int(irng, c_int)
      call check(c_mcica_subcol_lw(int(iplon,c_int), int(ncol,c_int), int(size(play,1),c_int), int(nlay,c_int), &
            int(icld,c_int), int(permuteseed,c_int), ir, play, cldfrac, ciwp, clwp, rei, rel, tauc, alpha, &
            cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl, RRTMG_LW_HOST, c_null_ptr), 'MCICA_SUBCOL_LW')
      irng = int(ir, im)
      end subroutine mcica_subcol_lw

      end module mcica_subcol_gen_lw
