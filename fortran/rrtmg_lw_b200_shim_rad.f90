! rrtmg_lw_b200_shim_rad.f90 -- module rrtmg_lw_rad of the McICA build (replaces src/rrtmg_lw_rad.f90): the
! reference's rrtmg_lw(ncol, nlay, icld, idrv, ..., cldfmcl, taucmcl, ...) argument for argument, routed to
! rrtmg_lw_b200_rrtmg_lw.  Compile after rrtmg_lw_b200_shim.f90 (module rrtmg_lw_b200_c); use EITHER this file OR
! rrtmg_lw_b200_shim_rad_nomcica.f90, as the reference ships two source files for the one module name.
! STATUS: not compiled in the development image (no Fortran compiler there).
!------------------------------------------------------------------------------------------------
      module rrtmg_lw_rad
      use iso_c_binding
      use parkind, only : im => kind_im, rb => kind_rb
      use rrtmg_lw_b200_c
      implicit none
      public :: rrtmg_lw, inatm

      contains

      subroutine rrtmg_lw &
            (ncol    ,nlay    ,icld    ,idrv    , &
             play    ,plev    ,tlay    ,tlev    ,tsfc    , &
             h2ovmr  ,o3vmr   ,co2vmr  ,ch4vmr  ,n2ovmr  ,o2vmr , &
             cfc11vmr,cfc12vmr,cfc22vmr,ccl4vmr ,emis    , &
             inflglw ,iceflglw,liqflglw,cldfmcl , &
             taucmcl ,ciwpmcl ,clwpmcl ,reicmcl ,relqmcl , &
             tauaer  , &
             uflx    ,dflx    ,hr      ,uflxc   ,dflxc,  hrc, &
             duflx_dt,duflxc_dt )
      integer(kind=im), intent(in) :: ncol, nlay
      integer(kind=im), intent(inout) :: icld
      integer(kind=im), intent(in) :: idrv
      real(kind=rb), intent(in), contiguous, target :: play(:,:), plev(:,:), tlay(:,:), tlev(:,:), tsfc(:)
      real(kind=rb), intent(in), contiguous, target :: h2ovmr(:,:), o3vmr(:,:), co2vmr(:,:), ch4vmr(:,:), &
                                                       n2ovmr(:,:), o2vmr(:,:), cfc11vmr(:,:), cfc12vmr(:,:), &
                                                       cfc22vmr(:,:), ccl4vmr(:,:), emis(:,:)
      integer(kind=im), intent(in) :: inflglw, iceflglw, liqflglw
      real(kind=rb), intent(in), contiguous, target :: cldfmcl(:,:,:), ciwpmcl(:,:,:), clwpmcl(:,:,:), &
                                                       reicmcl(:,:), relqmcl(:,:), taucmcl(:,:,:), tauaer(:,:,:)
      real(kind=rb), intent(out), contiguous, target :: uflx(:,:), dflx(:,:), hr(:,:), uflxc(:,:), dflxc(:,:), hrc(:,:)
      real(kind=rb), intent(out), contiguous, target, optional :: duflx_dt(:,:), duflxc_dt(:,:)
      integer(c_int) :: ld, ic
      type(c_ptr) :: p_d, p_dc
      ! `contiguous` makes the compiler pack a strided actual argument; every array must then share
      ! the column extent, which is what a GCM chunk (pcols) gives
      ld = int(size(play, 1), c_int)
      if (size(plev,1) /= ld .or. size(tlay,1) /= ld .or. size(tlev,1) /= ld .or. size(h2ovmr,1) /= ld .or. &
          size(emis,1) /= ld .or. size(tauaer,1) /= ld .or. size(uflx,1) /= ld .or. size(hr,1) /= ld .or. &
          size(cldfmcl,2) /= ld .or. size(reicmcl,1) /= ld) then
         write(*,*) 'RRTMG_LW: arrays must share one column extent (size(x,1))'
         error stop 1
      end if
      p_d = c_null_ptr; p_dc = c_null_ptr
      if (present(duflx_dt)) p_d = c_loc(duflx_dt)
      if (present(duflxc_dt)) p_dc = c_loc(duflxc_dt)
      ic = int(icld, c_int)
      call check(c_rrtmg_lw(int(ncol,c_int), ld, int(nlay,c_int), ic, int(idrv,c_int), play, plev, tlay, tlev, &
            tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, &
            int(inflglw,c_int), int(iceflglw,c_int), int(liqflglw,c_int), cldfmcl, taucmcl, ciwpmcl, clwpmcl, &
            reicmcl, relqmcl, tauaer, uflx, dflx, hr, uflxc, dflxc, hrc, p_d, p_dc, RRTMG_LW_HOST, c_null_ptr), &
            'RRTMG_LW')
      icld = int(ic, im)      ! clamped to 2 when outside [0,3] (src/rrtmg_lw_rad.f90:470)
      end subroutine rrtmg_lw

      ! src/rrtmg_lw_rad.f90:598-605 -- name kept, see rrtmg_lw_b200_c::internal_only
      subroutine inatm (iplon, nlay, icld, iaer, &
              play, plev, tlay, tlev, tsfc, h2ovmr, &
              o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, &
              cfc22vmr, ccl4vmr, emis, inflglw, iceflglw, liqflglw, &
              cldfmcl, taucmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, tauaer, &
              nlayers, pavel, pz, tavel, tz, tbound, semiss, coldry, &
              wkl, wbrodl, wx, pwvcm, inflag, iceflag, liqflag, &
              cldfmc, taucmc, ciwpmc, clwpmc, reicmc, relqmc, taua)
      integer(kind=im), intent(in) :: iplon, nlay, icld, iaer, inflglw, iceflglw, liqflglw
      real(kind=rb), intent(in) :: play(:,:), plev(:,:), tlay(:,:), tlev(:,:), tsfc(:), h2ovmr(:,:), o3vmr(:,:), &
            co2vmr(:,:), ch4vmr(:,:), n2ovmr(:,:), o2vmr(:,:), cfc11vmr(:,:), cfc12vmr(:,:), cfc22vmr(:,:), &
            ccl4vmr(:,:), emis(:,:), cldfmcl(:,:,:), taucmcl(:,:,:), ciwpmcl(:,:,:), clwpmcl(:,:,:), reicmcl(:,:), &
            relqmcl(:,:), tauaer(:,:,:)
      integer(kind=im), intent(out) :: nlayers, inflag, iceflag, liqflag
      real(kind=rb), intent(out) :: pavel(:), pz(0:), tavel(:), tz(0:), tbound, semiss(:), coldry(:), wkl(:,:), &
            wbrodl(:), wx(:,:), pwvcm, cldfmc(:,:), taucmc(:,:), ciwpmc(:,:), clwpmc(:,:), reicmc(:), relqmc(:), taua(:,:)
      call internal_only('INATM')
      end subroutine inatm

      end module rrtmg_lw_rad

