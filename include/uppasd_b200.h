/*
 * uppasd_b200.h -- C ABI of libuppasd_b200.so, the B200-native replacement for UppASD's per-time-step
 * spin-dynamics hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Two layers are exported:
 *
 *  (1) LEGACY BOUNDARY -- the exact symbols the reference's Fortran host already calls when built with
 *      -DCUDA (reference: source/chelper.f90:166-186, source/sd_driver.f90:1118-1153).  They replace
 *      source/gpu_files/fortranData.cpp:141-185 and source/gpu_files/fort_helper.cpp:18-64 one for one:
 *      gfortran implicit-interface calling convention, every argument by reference, 4-byte integers,
 *      column-major arrays, 1-based atom indices.
 *
 *  (2) EXPLICIT API (asd_*) -- the same engine with every input passed explicitly, for the parts of the
 *      hot path the reference never put behind its native boundary: solver choice (SDEalgh 1 midpoint /
 *      5 Depondt), per-site damping / temperature / Lande arrays, biquadratic tables, Monte Carlo sweeps
 *      (mc_evolve, source/MonteCarlo/montecarlo.f90:44), effective field + energy
 *      (source/Hamiltonian/hamiltonianactions.f90:108), on-device observables
 *      (source/Measurement/prn_averages.f90:414-456) and on-device table construction
 *      (source/Hamiltonian/neighbourmap.f90:32, hamiltonianinit.f90:985).  INTEGRATION.md shows the
 *      iso_c_binding interface blocks a maintainer adds to sd_driver.f90 / mc_driver.f90.
 *
 * All asd_* functions return 0 on success and a negative code on failure; asd_last_error() returns the
 * message.  There is no CPU fallback: without a CUDA device every compute entry fails loudly.
 */
#ifndef UPPASD_B200_H
#define UPPASD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * (1) Legacy boundary (drop-in for source/gpu_files)
 * ---------------------------------------------------------------------------------------------- */

/* replaces fortranData.cpp:141-148; argument order fixed by chelper.f90:171-173.  Stores pointers only. */
void fortrandata_setconstants_(char* stt, int* SDEalgh, unsigned int* rstep, unsigned int* nstep,
                               unsigned int* Natom, unsigned int* Mensemble, unsigned int* max_no_neigh,
                               double* delta_t, double* gamma, double* k_bolt, double* mub, double* damping,
                               double* binderc, double* mavg, int* mompar, char* initexc, unsigned int* do_dm,
                               unsigned int* max_no_dmneigh, unsigned int* do_jtensor,
                               unsigned int* do_anisotropy, unsigned int* nHam);

/* replaces fortranData.cpp:151-180; argument order fixed by chelper.f90:181-184.  Stores pointers only. */
void fortrandata_setmatrices_(double* ncoup, unsigned int* nlist, unsigned int* nlistsize, double* beff,
                              double* b2eff, double* emomM, double* emom, double* emom2,
                              double* external_field, double* mmom, double* btorque, double* Temp_array,
                              double* mmom0, double* mmom2, double* mmomi, double* dm_vect,
                              unsigned int* dmlist, unsigned int* dmlistsize, double* j_tens, double* kaniso,
                              double* eaniso, unsigned int* taniso, double* sb, unsigned int* aHam);

/* replaces fortranData.cpp:183-185 (chelper.f90:186). */
void fortrandata_setinputdata_(int* gpu_mode, int* gpu_rng, int* gpu_rng_seed);

/* replace fort_helper.cpp:52-64 (called from sd_driver.f90:1140-1142).
 * cudamdsim_measurementphase_ samples asynchronously: the state of a sampled step is staged on the device, copied to pinned host
 * memory on a second stream and handed to fortran_measure_moment by a worker thread while the time loop continues (three
 * snapshots in flight; gpu_files/cudaMeasurement.cu:109-182 and measurementQueue.cpp:63-121 in the reference); the status line
 * takes Mbar from the on-device sums.  ASD_LEGACY_SYNC=1 selects blocking copies into the module arrays instead.
 * asd_legacy_async_samples(): samples served by the asynchronous path so far. */
void cudamdsim_initiateconstants_(void);
void cudamdsim_initiatematrices_(void);
void cudamdsim_measurementphase_(void);
long asd_legacy_async_samples(void);

/* gpu_mode 2 entry points (fort_helper.cpp:18-45, sd_driver.f90:1145-1147).  This build has no CPU twin:
 * they run the same CUDA engine. */
void cmdsim_initiateconstants_(void);
void cmdsim_initiatefortran_(void);
void cmdsim_measurementphase_(void);

/* New sibling entries, same calling convention (every argument by reference), for the two driver loops the
 * reference never moved behind its native boundary.  Both act on the engine that cudamdsim_initiatematrices_
 * built from the fortrandata_* pointers and write emom / emomM / mmom back to the Fortran arrays on return.
 *   cudamdsim_initialphase_  one phase of sd_iphase (sd_driver.f90:144-289): ipnstep(i), ipTemp(i), ipdelta_t(i),
 *                            iplambda1(i), ipSDEalgh; first_step keys the noise.
 *   cudamcsim_evolve_        nsweeps x mc_evolve (montecarlo.f90:44; call sites mc_driver.f90:126-132, 365-371, 518-524):
 *                            mode 'M' / 'H', Temp, temprescale, extfield(3); *upload != 0 re-reads emom / mmom first. */
void cudamdsim_initialphase_(unsigned int* ipnstep, double* ipTemp, double* ipdelta_t, double* iplambda1,
                             int* ipSDEalgh, unsigned int* first_step);
void cudamcsim_evolve_(char* mode, unsigned int* nsweeps, unsigned int* first_sweep, double* Temp,
                       double* temprescale, double* extfield, int* upload);

/* pyasd, the reference's second caller (source/pyasd.f90; the Python package uppasd binds these names): the same bind(c)
 * entry points, acting on the engine and the module arrays handed over by fortrandata_set*_ / cudamdsim_initiatematrices_.
 *   relax_       pyasd.f90:255-298  imode 'M' / 'H': instep sweeps of mc_evolve at itemperature (mc_minimal); otherwise
 *                                   instep midpoint steps at itemperature with damping idamping (sd_minimal, solver 1, the
 *                                   module's delta_t; itimestep is accepted and unused, as in the reference).
 *                                   moments(3,natom,mensemble) = emomM on return.
 *   get_emom_    pyasd.f90:316-328  moments = emom
 *   put_emom_    pyasd.f90:330-350  emom = moments; emom2 = emom; emomM = moments * mmom
 *   get_beff_    pyasd.f90:356-369  effective_field(); fields = beff
 *   get_energy_  pyasd.f90:505-517  effective_field(energy); energy / (Natom * Mensemble), in mRy */
void relax_(double* moments, int* natom, int* mensemble, char* imode, int* instep, double* itemperature,
            double* itimestep, double* idamping);
void get_emom_(double* moments, int* natom, int* mensemble);
void put_emom_(const double* moments, int* natom, int* mensemble);
void get_beff_(double* fields, int* natom, int* mensemble);
void get_energy_(double* energy);

/* Extra inputs the reference never passes through fortrandata_* but whose Fortran semantics the engine
 * honours when given (all optional; NULL keeps the legacy behaviour of one global damping, g=2):
 *   Landeg(N), lambda1_array(N) (evolution.f90:38-44), bqlist/j_bq/bqlistsize (hamiltoniandatatype.f90:58-61). */
void fortrandata_setextras_(double* Landeg, double* lambda1_array, double* temprescale, unsigned int* do_bq,
                            unsigned int* nn_bq_tot, unsigned int* bqlist, unsigned int* bqlistsize,
                            double* j_bq);

/* Shape of the supercell behind the tables of fortrandata_setmatrices_ (inputdata: NA, N1, N2, N3, BC1..BC3; the atom order
 * is i0 + NA*(ix + N1*(iy + N2*iz)), geometry.f90:440-460).  Optional: with it the engine orders the atoms in bricks and
 * the Fortran-built nlist runs on the fast path of device-built lattices (asd_set_lattice_hint); results do not change. */
void fortrandata_setlattice_(unsigned int* NA, unsigned int* N1, unsigned int* N2, unsigned int* N3, char* BC1, char* BC2,
                             char* BC3);

/* Callbacks into the host (reference: source/gpu_files/c_helper.h:30-37, bodies in chelper.f90:73-160).
 * When the library is linked into the Fortran program the gfortran-mangled symbols
 * __chelper_MOD_fortran_* are picked up automatically (weak references).  A non-Fortran host registers
 * them here instead.  mstep is passed as size_t* like the reference does (c_helper.h:30-36). */
typedef void (*asd_cb_do_measurements)(const size_t* mstep, int* do_copy);
typedef void (*asd_cb_measure_moment)(const double* emomM, const double* emom, const double* mmom,
                                      const size_t* mstep);
typedef void (*asd_cb_flush_measurements)(const size_t* mstep);
typedef void (*asd_cb_status)(double* mavg);
void asd_set_callbacks(asd_cb_do_measurements do_meas, asd_cb_measure_moment measure,
                       asd_cb_flush_measurements flush, asd_cb_status status);

/* ------------------------------------------------------------------------------------------------
 * (2) Explicit API
 * ---------------------------------------------------------------------------------------------- */
typedef struct asd_engine asd_engine;

/* the engine behind the legacy symbols (NULL before cudamdsim_initiateconstants_): lets a host use the explicit
 * API (asd_measure, asd_layout_info, ...) on the state the legacy calls created */
asd_engine* asd_legacy_engine(void);

const char* asd_last_error(void);
int asd_device_count(void);

/* device < 0: current device.  Fails when no CUDA device is present. */
int asd_create(asd_engine** out, int device);
void asd_destroy(asd_engine* e);

/* physical constants, passed like the reference does (constants.f90:14-29 are mutable; chelper.f90:171). */
int asd_set_constants(asd_engine* e, double gamma, double k_bolt, double mub, double mry);

/* sizes + Hamiltonian look-up table aHam(N) (hamiltonianinit.f90:818-838); nHam==Natom => aHam may be NULL. */
int asd_set_system(asd_engine* e, int Natom, int Mensemble, int nHam, const int* aHam);

/* Heisenberg table: nlist(z,N) 1-based, nlistsize(NH), ncoup(z,NH) (hamiltoniandatatype.f90:14-40). */
int asd_set_exchange(asd_engine* e, int max_no_neigh, const int* nlist, const int* nlistsize,
                     const double* ncoup);
/* DM table: dmlist(zdm,N), dmlistsize(NH), dm_vect(3,zdm,NH). */
/* Tensorial exchange instead of the scalar table (do_jtensor 1): j_tens(3,3,max_no_neigh,nHam) as mounted by
 * setup_neighbour_hamiltonian with hdim 9 (hamiltonianinit.f90:412-432); the field is tensor_field's
 * f += J(:,1) m_x + J(:,2) m_y + J(:,3) m_z (hamiltonianactions.f90:499-542), indexed by the Hamiltonian row aHam(i). */
int asd_set_jtensor(asd_engine* e, int max_no_neigh, const int* nlist, const int* nlistsize, const double* j_tens);
int asd_set_dm(asd_engine* e, int max_no_dmneigh, const int* dmlist, const int* dmlistsize,
               const double* dm_vect);
/* biquadratic table: bqlist(zbq,N), bqlistsize(NH), j_bq(zbq,NH). */
int asd_set_bq(asd_engine* e, int nn_bq_tot, const int* bqlist, const int* bqlistsize, const double* j_bq);
/* single-ion anisotropy: taniso(N) in {0,1,2,7}, eaniso(3,N), kaniso(2,N), sb(N). */
/* Optional: the tables describe a supercell of N1 x N2 x N3 cells with NA atoms each in the reference's atom order
 * (i0 + NA*(ix + N1*(iy + N2*iz))); bc3 = "PP0"-style boundary conditions.  The engine then stores the atoms in brick
 * order (like asd_build_lattice_table does) and checks ON THE DEVICE whether the host's tables have the regularity the
 * run-compressed kernel needs; if not, or without this call, the one-atom-per-thread kernels run.  Results are
 * independent of the hint (parity bar 1e-12). */
int asd_set_lattice_hint(asd_engine* e, int NA, int N1, int N2, int N3, const char* bc3);

int asd_set_anisotropy(asd_engine* e, const int* taniso, const double* eaniso, const double* kaniso,
                       const double* sb);
/* external_field(3,N,M) (calculatefields.f90:23-82) and optional spin-transfer torque field btorque(3,N,M). */
int asd_set_external_field(asd_engine* e, const double* external_field);
int asd_set_torque(asd_engine* e, const double* btorque);

/* Time-dependent uniform field: the global part of calc_external_time_fields (magnetic-field pulse, microwave field;
 * calculatefields.f90:92-185), which hamiltonianactions.f90:241 adds to beff2 next to external_field.  tfield(3, nsteps) holds the
 * field of the steps first_step .. first_step + nsteps - 1 (the value of mstep the step is run with), the same for every
 * ensemble (the one ensemble-dependent term of the reference, the demagnetisation field, is outside this path).  The stage
 * kernels take the vector of their step as a kernel parameter -- no re-upload of the per-site external_field between steps.
 * Steps outside the schedule see no time-dependent field; NULL or nsteps <= 0 clears it.  LLG steps only (the reference's Monte
 * Carlo takes its field as the extfield argument of mc_evolve). */
int asd_set_time_field(asd_engine* e, long first_step, long nsteps, const double* tfield);

/* LLG parameters (evolution.f90:38-44): SDEalgh 1 (midpoint) or 5 (Depondt); per-site arrays of length N.
 * mompar as updatemoments.f90:105-145; seed keys the counter-based noise generator. */
int asd_set_llg(asd_engine* e, int SDEalgh, double delta_t, const double* Landeg, const double* lambda1_array,
                const double* Temp_array, double temprescale, int mompar, unsigned long long seed);

/* moments: emom(3,N,M) unit vectors, mmom(N,M) magnitudes, mmom0(N,M) (NULL => mmom). */
/* Fixed-moment runs: red_atom_list(Nred) = the 1-based atoms that evolve, as evolve_first / evolve_second receive it
 * (evolution.f90:38-44, midpoint.f90:123, depondt.f90:138); every other atom keeps its moment and still acts on its
 * neighbours.  Nred <= 0 or a null list: every atom evolves.  Monte Carlo sweeps ignore the list, as mc_evolve does.
 * Large systems step with the direct one-atom-per-thread kernel while a list is set (the tile kernels carry no mask). */
int asd_set_evolving_atoms(asd_engine* e, int Nred, const int* red_atom_list);
int asd_set_moments(asd_engine* e, const double* emom, const double* mmom, const double* mmom0);
int asd_get_moments(asd_engine* e, double* emom, double* emomM, double* mmom);

/* Freeze the tables into the device layout.  Must be called after the set_* calls and before compute. */
int asd_commit(asd_engine* e);

/* effective_field (hamiltonianactions.f90:108-252): beff(3,N,M) [, beff1, beff2] on the host (any may be
 * NULL), energy[M] in mRy with the reference's estimator (:245-250), summed over atoms per ensemble. */
int asd_effective_field(asd_engine* e, double* beff, double* beff1, double* beff2, double* energy);

/* nsteps LLG steps = field / evolve_first / field / evolve_second / moment_update each
 * (sd_driver.f90:668-764).  first_step is the value of mstep for the first step (keys the noise). */
int asd_sd_steps(asd_engine* e, long nsteps, long first_step);

/* The measurement-phase loop (sd_mphase, sd_driver.f90:517-849) in one call: nsteps LLG steps, and after every
 * `sample_every`-th step the per-ensemble sums of emomM that prn_averages buffers (prn_averages.f90:437-447) are reduced
 * on the device and kept in a sample ring.  One device-to-host copy and one synchronisation end the call:
 * msum(3, Mensemble, nsamples) with nsamples = nsteps / sample_every (also returned in *nsamples; msum may be NULL to
 * leave the samples on the device).  Replaces the per-sample full-state copy of the reference's CUDA loop
 * (gpu_files/cudaMdSimulation.cu:400-470, cudaMeasurement.cu:109-182).  sample_every <= 0: one sample after the last step. */
int asd_sd_run(asd_engine* e, long nsteps, long first_step, long sample_every, double* msum, long* nsamples);

/* nsweeps Monte Carlo sweeps (mc_evolve, montecarlo.f90:44-273): mode 'M' Metropolis / 'H' heat bath,
 * N*M single-spin trials per sweep visited colour by colour (graph colouring of the union of all
 * neighbour tables).  extfield[3] is mc_evolve's uniform field argument. */
int asd_mc_sweeps(asd_engine* e, char mode, long nsweeps, long first_sweep, double temperature,
                  double temprescale, const double* extfield);

/* Colouring used by asd_mc_sweeps.  layout 0: colour-major re-ordering of the atoms (greedy colouring of the actual
 * neighbour graph; default for undecomposed engines), layout 1: lattice (brick) order with a PERIODIC colouring
 * (colour = f(basis atom, global cell coordinates mod period)) -- always used by a slab, where it gives one halo
 * exchange per colour and the same Markov chain for every decomposition; selectable on any device-built lattice.
 * asd_mc_colouring reports what the next sweep uses: number of colours and (layout 1) the period in cells. */
int asd_set_mc_layout(asd_engine* e, int layout);   /* -1 automatic, 0, 1, 2 (block sweep, below) */
int asd_mc_colouring(asd_engine* e, int* layout, int* ncolours, int* period3);
/* colour (0-based) of every atom of this engine, original atom order [Natom] */
int asd_get_mc_colours(asd_engine* e, int* colour);
/* Layout 2 (asd_set_mc_layout(e, 2); the default of large undecomposed device-built lattices): BLOCK SWEEP -- tiles of the
 * brick order are coloured as well, one CTA sweeps all atom colours of its tile in shared memory (asd_mc_block.cuh).
 * order[Natom]: the 1-based atoms in a sequential visiting order that reproduces the chain of the next sweep for every
 * layout (mc_evolve's iflip_a, montecarlo.f90:165-173): colour-parallel updates commute inside a colour class. */
int asd_get_mc_visit_order(asd_engine* e, int* order);
/* Test hook: the random draws of Monte Carlo sweep `sweep` exactly as the update kernels compute them from the counter-based
 * generator: u(4,Natom,Mensemble) uniforms (Metropolis: move type, azimuth, cos(theta), acceptance; heat bath: polar draw,
 * azimuth), g(3,Natom,Mensemble) the Gaussian trial-move numbers.  Lets a CPU restatement of mc_evolve replay a sweep. */
int asd_debug_mc_draws(asd_engine* e, long sweep, double* u, double* g);

/* On-device observables: msum(3,M) = sum_i emomM(:,i,k) (prn_averages.f90:437-447); energy[M] as above
 * (NULL to skip). */
int asd_measure(asd_engine* e, double* msum, double* energy);

/* Term-resolved energy per atom in mRy (calc_energy, energy.f90:180-340, the columns of totenergy.*.out that
 * exist on this path): terms(5,M) = exchange, anisotropy, DM, biquadratic, Zeeman; their sum is ene%energy. */
int asd_energy_terms(asd_engine* e, double* terms);

/* Sublattice-projected sums (buffer_proj_avrg, prn_averages.f90:462-512): msum_na(3,NA,M) = sum of emomM over the atoms
 * with basis number mod(i-1,NA)+1 = i_na.  projavgs.*.out divides by N1*N2*N3 and, for do_proj_avrg Y, adds the basis
 * atoms of one type (prn_proj_avrg, :662-760). */
int asd_measure_sublattice(asd_engine* e, int NA, double* msum_na);

/* Skyrmion number by triangulation (skyno T): simp(3,nsimp) = 1-based atom numbers of the triangle corners as
 * delaunay_tri_tri builds them (topology.f90:307-380); q[M] = sum over triangles of the signed solid angle / 4 pi per
 * ensemble (pontryagin_tri, topology.f90:78-116, before its division by Mensemble; buffer_skyno_tri divides by NA). */
int asd_set_triangulation(asd_engine* e, int nsimp, const int* simp);
int asd_skyrmion_number(asd_engine* e, double* q);

/* selected moments for trajectory output (prn_trajectories.f90:60-110): atoms[n] 1-based, out(4,n,M) = ex,ey,ez,|m| */
int asd_get_atoms(asd_engine* e, int n, const int* atoms, double* out);

/* Device-timing helper for bench.py: runs nsteps steps bracketed by CUDA events on the engine's stream
 * and returns the elapsed milliseconds; per-kernel time of the two stage kernels in stage_ms[2]. */
int asd_time_sd_steps(asd_engine* e, long nsteps, long first_step, float* total_ms, float* stage_ms);
int asd_time_mc_sweeps(asd_engine* e, char mode, long nsweeps, double temperature, float* total_ms);

/* number of kernel launches issued by this engine so far (bench.py's gpu_launches). */
/* which field path the LLG stage kernels of the committed layout use: info[0] = 1 if the tile's gather list is staged
 * in shared memory, info[1] = R of the run-compressed register-blocked kernel (0: one atom per thread), info[2] =
 * largest gather list of a tile, info[3] = largest number of distinct neighbour runs of a group of R runs,
 * info[4] = slots per tile (256, 512 or 1024), info[5] = bit 0: the DM / BQ neighbours are read from shared memory too,
 * bit 1: the gather list is staged from the moment planes emomM[M][3][Npad] with asynchronous copies (MM instantiations) */
int asd_layout_info(asd_engine* e, int* info6);
long asd_launch_count(asd_engine* e);
int asd_synchronize(asd_engine* e);

/* ---- on-device table construction ----------------------------------------------
 * Builds nlist / couplings for a periodic or open supercell directly in device memory from the unit-cell
 * stencil, bit-identical to setup_nm + setup_neighbour_hamiltonian (neighbourmap.f90:248-321,
 * hamiltonianinit.f90:1040-1091).  kind: 0 exchange, 1 DM (3 components), 2 biquadratic.
 *   nslot[NA]                 number of stencil entries of basis atom i0
 *   cell_atom[NA*maxslot]     j0 (1-based basis atom hit), stencil order = shell-major, image order
 *   cell_shift[NA*maxslot*3]  (dx,dy,dz) cell translation of each entry
 *   coupling[NA*maxslot*ncomp] coupling of each entry, already unit-converted (what ncoup would hold)
 * The engine must have been given asd_set_system with do_reduced semantics (nHam = NA or Natom). */
int asd_build_lattice_table(asd_engine* e, int kind, int NA, int N1, int N2, int N3, const char* bc3,
                            int maxslot, const int* nslot, const int* cell_atom, const int* cell_shift,
                            const double* coupling);
/* copies a built / uploaded table back in Fortran layout: list(z,N) 1-based, listsize(NH), coup(ncomp,z,NH). */
int asd_get_table_dims(asd_engine* e, int kind, int* z, int* ncomp);
int asd_get_table(asd_engine* e, int kind, int* list, int* listsize, double* coup);

/* moments generated on the device for large synthetic runs (bench): e_i = normalize(1, a sin(2 pi h_i),
 * a cos(2 pi h_i)), h_i = frac(i * 0.6180339887), i the 1-based atom index; magnitude per basis atom. */
int asd_init_moments_tilted(asd_engine* e, double amplitude, int NA, const double* mmom_basis);

/* ---- multi-GPU (neither mode exists in the reference, which is single-device) ------------------
 * Ensemble sharding: every engine holds Mensemble/G whole ensembles and all tables; no communication.  The
 * noise is keyed by the GLOBAL ensemble index so that results do not depend on G. */
int asd_set_ensemble_offset(asd_engine* e, unsigned int first_ensemble);

/* Slab decomposition of ONE supercell along z: engine `slab_index` of `nslabs` owns N3/nslabs consecutive cell
 * planes (the atom index i0 + NA*(ix + N1*(iy + N2*iz)) of geometry.f90 makes that a contiguous index range).
 * Call after asd_set_system (Natom = atoms of the local slab) and before asd_build_lattice_table, which then
 * takes the GLOBAL N3.  halo_planes = interaction range along z in cell planes (2 for the 4-shell bcc table).
 * Neighbour entries that leave the slab point at halo slots; the stage kernels of the boundary tiles store
 * their new spins directly into the ring neighbours' halo slots (peer memory over NVLink) and publish an epoch
 * flag; the next stage waits on the flags.  Noise and tilted initial moments are keyed by the global atom index.
 * Monte Carlo sweeps on a slab visit the atoms by a periodic colouring that every slab derives from the global cell
 * coordinates: one halo exchange per colour, fused into the boundary-tile launches in the same way.
 * Arrays passed to / returned by asd_set_moments / asd_get_moments / asd_measure are those of the local slab. */
int asd_set_slab(asd_engine* e, int nslabs, int slab_index, int halo_planes);
/* after asd_commit: IPC handles (asd_slab_handle_bytes() bytes) of this slab's cur / pred / flag buffers and its two moment-plane buffers ... */
int asd_slab_handle_bytes(void);
int asd_slab_export(asd_engine* e, void* handles);
/* ... which the ring neighbours open (one process per GPU; exchange the bytes with any host transport) */
int asd_slab_connect_ipc(asd_engine* e, const void* lower_handles, const void* upper_handles);
/* same, for engines that live in one process (also: a single slab whose halos are its own periodic images) */
int asd_slab_connect_local(asd_engine* e, asd_engine* lower, asd_engine* upper);
/* exchanges completed so far; error_flag != 0 (and a negative return) if a halo wait timed out */
int asd_slab_status(asd_engine* e, unsigned long long* epoch, int* error_flag);

#ifdef __cplusplus
}
#endif
#endif /* UPPASD_B200_H */
