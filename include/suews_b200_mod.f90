! suews_b200_mod.f90 -- ISO_C_BINDING interfaces to libsuews_b200.so for Fortran hosts (the SUEWS binary's grid
! loop, suews_program.f95:606-623; WRF-SUEWS' SuMin, suews_ctrl_sumin.f95:543).  One interface per function of
! include/suews_b200.h; every function returns 0 or a SUEWS error code (suews_last_error holds the text).
!
! The reference-shaped entry points need no interface module: SUEWS_cal_Main / SUEWS_cal_multitsteps are exported under
! the external-procedure symbols gfortran generates (suews_cal_main_, suews_cal_multitsteps_; INTEGRATION.md section 2),
! so an EXTERNAL declaration -- or the explicit interface of INTEGRATION.md -- binds them.
!
! Not compiled in the authoring image (it has no Fortran compiler); written against Fortran 2003.
!
! Typical use (all grids of a run, one block of timesteps per call):
!
!    USE suews_b200
!    TYPE(C_PTR) :: h
!    ierr = suews_b200_create(NumberOfGrids, 1_C_INT, 0_C_INT, h)
!    ierr = suews_b200_set_class_param(h, 0_C_INT, 'NetRadiationMethod'//C_NULL_CHAR, [REAL(NetRadiationMethod, C_DOUBLE)], 1_C_INT)
!    ierr = suews_b200_set_cell_param(h, 'sfr'//C_NULL_CHAR, sfr_all, 7_C_INT)          ! sfr_all(7, NumberOfGrids)
!    ierr = suews_b200_set_state(h, 'state_id'//C_NULL_CHAR, state_all, 7_C_INT)
!    ierr = suews_b200_upload(h)
!    ierr = suews_b200_set_forcing(h, MetForcingData(:, 1:24, 1), irMax)                ! shared forcing, MultipleMetFiles = 0
!    ierr = suews_b200_advance(h, 0_C_INT, irMax, SUEWS_OUT_AGG)
!    ierr = suews_b200_sync(h)
!    ierr = suews_b200_get_agg_block(h, igrid - 1, dataOut_agg, irMax/nlinesOut)        ! what SUEWS_Write_txt formats
!    ierr = suews_b200_get_state(h, 'state_id'//C_NULL_CHAR, state_all, 7_C_INT)
MODULE suews_b200
   USE ISO_C_BINDING
   IMPLICIT NONE

   INTEGER(C_INT), PARAMETER :: SUEWS_OK = 0
   INTEGER(C_INT), PARAMETER :: SUEWS_E_NO_DEVICE = 2001, SUEWS_E_BAD_NAME = 2002, SUEWS_E_BAD_SIZE = 2003, &
                                SUEWS_E_BAD_ARG = 2004, SUEWS_E_NOT_IMPL = 1003, SUEWS_E_ESTM_INPUT = 1005
   ! output policy bits of suews_b200_advance
   INTEGER(C_INT), PARAMETER :: SUEWS_OUT_SUM82 = 1, SUEWS_OUT_FULL = 2, SUEWS_OUT_SUM7 = 4, SUEWS_OUT_AGG = 8, &
                                SUEWS_OUT_FLUX_ONLY = 16

   INTERFACE
      INTEGER(C_INT) FUNCTION suews_b200_device_count() BIND(C, name='suews_b200_device_count')
         IMPORT :: C_INT
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_create(ncell, nclass, device, h) BIND(C, name='suews_b200_create')
         IMPORT :: C_INT, C_PTR
         INTEGER(C_INT), VALUE :: ncell, nclass, device
         TYPE(C_PTR), INTENT(OUT) :: h
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_destroy(h) BIND(C, name='suews_b200_destroy')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
      END FUNCTION

      ! name: the reference's argument name, C_NULL_CHAR terminated; kind: 0 class int, 1 class real, 2 per cell,
      ! 3 state, 4 ESTM state; returns the element count or -1
      INTEGER(C_INT) FUNCTION suews_b200_field_info(name, kind) BIND(C, name='suews_b200_field_info')
         IMPORT :: C_INT, C_CHAR
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         INTEGER(C_INT), INTENT(OUT) :: kind
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_class_param(h, cls, name, vals, n) BIND(C, name='suews_b200_set_class_param')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: cls, n
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
      END FUNCTION

      ! vals(n, ncell): the n values of cell 1, then of cell 2, ...
      INTEGER(C_INT) FUNCTION suews_b200_set_cell_param(h, name, vals, n) BIND(C, name='suews_b200_set_cell_param')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_cell_class(h, class_of_cell) BIND(C, name='suews_b200_set_cell_class')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), INTENT(IN) :: class_of_cell(*)      ! 0-based class index per cell
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_state(h, name, vals, n) BIND(C, name='suews_b200_set_state')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_get_state(h, name, vals, n) BIND(C, name='suews_b200_get_state')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(OUT) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_clock(h, dt_since_start, tstep_prev) BIND(C, name='suews_b200_set_clock')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: dt_since_start, tstep_prev
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_get_clock(h, dt_since_start, tstep_prev) BIND(C, name='suews_b200_get_clock')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), INTENT(OUT) :: dt_since_start, tstep_prev
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_upload(h) BIND(C, name='suews_b200_upload')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
      END FUNCTION

      ! met(nsteps, 24): MetForcingData(:, 1:24, igrid), the columns of suews_ctrl_driver.f95:3177-3201
      INTEGER(C_INT) FUNCTION suews_b200_set_forcing(h, met, nsteps) BIND(C, name='suews_b200_set_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: met(*)
         INTEGER(C_INT), VALUE :: nsteps
      END FUNCTION

      ! met_orig(nrows, 24): the file's rows (MetRead); disaggregated to the model timestep on the device
      INTEGER(C_INT) FUNCTION suews_b200_set_forcing_orig(h, met_orig, nrows, resolution_in_s, disagg_method, &
                                                         rain_disagg_method, raw_units, smd_method) &
         BIND(C, name='suews_b200_set_forcing_orig')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: met_orig(*)
         INTEGER(C_INT), VALUE :: nrows, resolution_in_s, disagg_method, rain_disagg_method, raw_units, smd_method
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_get_forcing_block(h, met_fast, nsteps) BIND(C, name='suews_b200_get_forcing_block')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(OUT) :: met_fast(*)
         INTEGER(C_INT), VALUE :: nsteps
      END FUNCTION

      ! ts5(nsteps, 13): ESTMForcingData(:, 1:13, igrid) (StorageHeatMethod 4/14)
      INTEGER(C_INT) FUNCTION suews_b200_set_estm_forcing(h, ts5, nsteps) BIND(C, name='suews_b200_set_estm_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: ts5(*)
         INTEGER(C_INT), VALUE :: nsteps
      END FUNCTION

      ! KdownZen = 1 on the block set_forcing_orig disaggregated (suews_ctrl_input.f95:1568-1600)
      INTEGER(C_INT) FUNCTION suews_b200_apply_kdown_zen(h, lat, lng, timezone, alt) BIND(C, name='suews_b200_apply_kdown_zen')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), VALUE :: lat, lng, timezone, alt
      END FUNCTION

      ! met(nrows, 24): MetForcingData(:, :, igrid), AnOHM's MetForcingData_grid when it is not the staged block (StorageHeatMethod 3)
      INTEGER(C_INT) FUNCTION suews_b200_set_anohm_forcing(h, met, nrows) BIND(C, name='suews_b200_set_anohm_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: met(*)
         INTEGER(C_INT), VALUE :: nrows
      END FUNCTION

      ! ini(8, nline): TRANSPOSE(IniCBLdata) -- the rows of CBL_initial_data.txt, one row per column here (BLUEWS, CBLuse >= 1)
      INTEGER(C_INT) FUNCTION suews_b200_set_cbl_initial(h, ini, nline) BIND(C, name='suews_b200_set_cbl_initial')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: ini(*)
         INTEGER(C_INT), VALUE :: nline
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_estm_forcing_orig(h, ts_orig, nrows, resolution_in_s, disagg_method) &
         BIND(C, name='suews_b200_set_estm_forcing_orig')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(IN) :: ts_orig(*)
         INTEGER(C_INT), VALUE :: nrows, resolution_in_s, disagg_method
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_advance(h, step0, nsteps, out_flags) BIND(C, name='suews_b200_advance')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: step0, nsteps, out_flags    ! step0 is 0-based
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_sync(h) BIND(C, name='suews_b200_sync')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_set_output_freq(h, out_freq_s) BIND(C, name='suews_b200_set_output_freq')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: out_freq_s
      END FUNCTION

      ! block(nrows, 87) of cell `cell` (0-based): what SUEWS_Output_Agg returns for that grid
      INTEGER(C_INT) FUNCTION suews_b200_get_agg_block(h, cell, block, nrows) BIND(C, name='suews_b200_get_agg_block')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: cell, nrows
         REAL(C_DOUBLE), INTENT(OUT) :: block(*)
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_agg_daily_rows(h, nrows) BIND(C, name='suews_b200_agg_daily_rows')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), INTENT(OUT) :: nrows
      END FUNCTION

      ! block(nrows, 50): the DailyState rows of the days that ended inside the last SUEWS_OUT_AGG launch
      INTEGER(C_INT) FUNCTION suews_b200_get_agg_daily_block(h, cell, block, nrows) BIND(C, name='suews_b200_get_agg_daily_block')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         INTEGER(C_INT), VALUE :: cell, nrows
         REAL(C_DOUBLE), INTENT(OUT) :: block(*)
      END FUNCTION

      ! sums(82, ncell)
      INTEGER(C_INT) FUNCTION suews_b200_get_sums(h, sums) BIND(C, name='suews_b200_get_sums')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         REAL(C_DOUBLE), INTENT(OUT) :: sums(*)
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_reset_sums(h) BIND(C, name='suews_b200_reset_sums')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
      END FUNCTION

      ! group: 'SUEWS', 'snow', 'RSL', 'DailyState', 'SOLWEIG', 'ESTM' (C_NULL_CHAR terminated); block(nsteps, ncol) as
      ! dataOutBlockX of SUEWS_cal_multitsteps (after SUEWS_OUT_FULL)
      INTEGER(C_INT) FUNCTION suews_b200_get_block(h, group, cell, block, nsteps, ncol) BIND(C, name='suews_b200_get_block')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: h
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: group(*)
         INTEGER(C_INT), VALUE :: cell, nsteps, ncol
         REAL(C_DOUBLE), INTENT(OUT) :: block(*)
      END FUNCTION

      ! sort the thread slots again on the current state (snow activity, canopy height); the run carries on
      INTEGER(C_INT) FUNCTION suews_b200_resort(h) BIND(C, name='suews_b200_resort')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: h
      END FUNCTION

      ! ---- several GPUs of one box behind one handle: contiguous blocks of grids, one per device, one host thread.
      !    ierr = suews_b200_multi_create(NumberOfGrids, 1_C_INT, ndev, devices, m)          ! devices(ndev) = 0, 1, ...
      !    ierr = suews_b200_multi_set_cell_param(m, 'sfr'//C_NULL_CHAR, sfr_all, 7_C_INT)   ! arrays of the WHOLE run
      !    ierr = suews_b200_multi_upload(m); ierr = suews_b200_multi_set_forcing(m, met, irMax)
      !    ierr = suews_b200_multi_advance(m, 0_C_INT, irMax, SUEWS_OUT_SUM7); ierr = suews_b200_multi_sync(m)
      !    ierr = suews_b200_multi_gather_sums(m, cols, 7_C_INT, sums, ms)                   ! sums(7, NumberOfGrids)
      INTEGER(C_INT) FUNCTION suews_b200_multi_create(ncell, nclass, ndev, devices, m) BIND(C, name='suews_b200_multi_create')
         IMPORT :: C_INT, C_PTR
         INTEGER(C_INT), VALUE :: ncell, nclass, ndev
         INTEGER(C_INT), INTENT(IN) :: devices(*)
         TYPE(C_PTR), INTENT(OUT) :: m
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_destroy(m) BIND(C, name='suews_b200_multi_destroy')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_ndev(m, ndev) BIND(C, name='suews_b200_multi_ndev')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), INTENT(OUT) :: ndev
      END FUNCTION

      ! block d (0-based) of the partition: its single-device handle, first grid (0-based) and grid count
      INTEGER(C_INT) FUNCTION suews_b200_multi_part(m, d, h, cell0, ncell) BIND(C, name='suews_b200_multi_part')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), VALUE :: d
         TYPE(C_PTR), INTENT(OUT) :: h
         INTEGER(C_INT), INTENT(OUT) :: cell0, ncell
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_class_param(m, cls, name, vals, n) BIND(C, name='suews_b200_multi_set_class_param')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), VALUE :: cls, n
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_cell_param(m, name, vals, n) BIND(C, name='suews_b200_multi_set_cell_param')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_cell_class(m, class_of_cell) BIND(C, name='suews_b200_multi_set_cell_class')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), INTENT(IN) :: class_of_cell(*)
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_state(m, name, vals, n) BIND(C, name='suews_b200_multi_set_state')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_get_state(m, name, vals, n) BIND(C, name='suews_b200_multi_get_state')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(OUT) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_clock(m, dt_since_start, tstep_prev) BIND(C, name='suews_b200_multi_set_clock')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), VALUE :: dt_since_start, tstep_prev
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_upload(m) BIND(C, name='suews_b200_multi_upload')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_forcing(m, met, nsteps) BIND(C, name='suews_b200_multi_set_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: met(*)
         INTEGER(C_INT), VALUE :: nsteps
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_forcing_orig(m, met_orig, nrows, resolution_in_s, disagg_method, &
                                                                rain_disagg_method, raw_units, smd_method) &
         BIND(C, name='suews_b200_multi_set_forcing_orig')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: met_orig(*)
         INTEGER(C_INT), VALUE :: nrows, resolution_in_s, disagg_method, rain_disagg_method, raw_units, smd_method
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_apply_kdown_zen(m, lat, lng, timezone, alt) BIND(C, name='suews_b200_multi_apply_kdown_zen')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), VALUE :: lat, lng, timezone, alt
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_estm_forcing(m, ts5, nsteps) BIND(C, name='suews_b200_multi_set_estm_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: ts5(*)
         INTEGER(C_INT), VALUE :: nsteps
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_estm_forcing_orig(m, ts_orig, nrows, resolution_in_s, disagg_method) &
         BIND(C, name='suews_b200_multi_set_estm_forcing_orig')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: ts_orig(*)
         INTEGER(C_INT), VALUE :: nrows, resolution_in_s, disagg_method
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_anohm_forcing(m, met, nrows) BIND(C, name='suews_b200_multi_set_anohm_forcing')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: met(*)
         INTEGER(C_INT), VALUE :: nrows
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_set_cbl_initial(m, ini, nline) BIND(C, name='suews_b200_multi_set_cbl_initial')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         REAL(C_DOUBLE), INTENT(IN) :: ini(*)
         INTEGER(C_INT), VALUE :: nline
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_get_block(m, group, cell, block, nsteps, ncol) BIND(C, name='suews_b200_multi_get_block')
         IMPORT :: C_INT, C_PTR, C_CHAR, C_DOUBLE
         TYPE(C_PTR), VALUE :: m
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: group(*)
         INTEGER(C_INT), VALUE :: cell, nsteps, ncol
         REAL(C_DOUBLE), INTENT(OUT) :: block(*)
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_advance(m, step0, nsteps, out_flags) BIND(C, name='suews_b200_multi_advance')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), VALUE :: step0, nsteps, out_flags
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_sync(m) BIND(C, name='suews_b200_multi_sync')
         IMPORT :: C_INT, C_PTR
         TYPE(C_PTR), VALUE :: m
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_b200_multi_last_advance_ms(m, ms) BIND(C, name='suews_b200_multi_last_advance_ms')
         IMPORT :: C_INT, C_PTR, C_FLOAT
         TYPE(C_PTR), VALUE :: m
         REAL(C_FLOAT), INTENT(OUT) :: ms
      END FUNCTION

      ! the final gather: out(ncols, ncell) -- the selected columns of every grid's running sums
      INTEGER(C_INT) FUNCTION suews_b200_multi_gather_sums(m, cols, ncols, out, gather_ms) BIND(C, name='suews_b200_multi_gather_sums')
         IMPORT :: C_INT, C_PTR, C_DOUBLE, C_FLOAT
         TYPE(C_PTR), VALUE :: m
         INTEGER(C_INT), INTENT(IN) :: cols(*)
         INTEGER(C_INT), VALUE :: ncols
         REAL(C_DOUBLE), INTENT(OUT) :: out(*)
         REAL(C_FLOAT), INTENT(OUT) :: gather_ms
      END FUNCTION

      ! MODULE ESTM_data for the reference-shaped calls (INTEGRATION.md, "Semantics worth knowing")
      ! MetForcingData_grid for the reference-shaped calls with StorageHeatMethod 3: met(nrows, 24)
      INTEGER(C_INT) FUNCTION suews_anohm_set_met_grid(met, nrows) BIND(C, name='suews_anohm_set_met_grid')
         IMPORT :: C_INT, C_DOUBLE
         REAL(C_DOUBLE), INTENT(IN) :: met(*)
         INTEGER(C_INT), VALUE :: nrows
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_estm_data_set(name, vals, n) BIND(C, name='suews_estm_data_set')
         IMPORT :: C_INT, C_CHAR, C_DOUBLE
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(IN) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_estm_data_get(name, vals, n) BIND(C, name='suews_estm_data_get')
         IMPORT :: C_INT, C_CHAR, C_DOUBLE
         CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
         REAL(C_DOUBLE), INTENT(OUT) :: vals(*)
         INTEGER(C_INT), VALUE :: n
      END FUNCTION

      ! replaces the ErrorHint -> problems.txt -> STOP convention: code and message of the calling thread's last error
      INTEGER(C_INT) FUNCTION suews_last_error(code, buf, len) BIND(C, name='suews_last_error')
         IMPORT :: C_INT, C_CHAR
         INTEGER(C_INT), INTENT(OUT) :: code
         CHARACTER(KIND=C_CHAR), INTENT(OUT) :: buf(*)
         INTEGER(C_INT), VALUE :: len
      END FUNCTION

      INTEGER(C_INT) FUNCTION suews_cal_sunposition_c(year, idectime, UTC, lat, lng, alt, sunazimuth, sunzenith) &
         BIND(C, name='suews_cal_sunposition_c')
         IMPORT :: C_INT, C_DOUBLE
         REAL(C_DOUBLE), VALUE :: year, idectime, UTC, lat, lng, alt
         REAL(C_DOUBLE), INTENT(OUT) :: sunazimuth, sunzenith
      END FUNCTION
   END INTERFACE
END MODULE suews_b200
