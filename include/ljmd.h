/*
 * ljmd.h — C ABI of the B200-native Lennard-Jones MD step.
 *
 * This is the drop-in boundary for the hot path of vlvovch/lennard-jones-cuda
 * (all citations are into /root/reference/src/library/):
 *
 *   - MDSystem.cpp:9-25 declares the `extern "C"` seam the reference host layer
 *     links against (defined in MDSystem.cu:155-299).  Section B below exports
 *     the same six symbols the host layer actually calls, so the UNMODIFIED
 *     reference MDSystem.cpp (built with -DUSE_CUDA_TOOLKIT) links against this
 *     library instead of MDSystem.cu.
 *   - Section A is the device-resident handle API that the source-compatible
 *     `MDSystem` class (lennard-jones-cuda_b200/host/MDSystem.h) is written on:
 *     one call per MDSystem method on the hot path, state kept in HBM between
 *     steps, error codes instead of exit(), optional sharding of the i-particles
 *     over several GPUs (one process per GPU).
 *
 * Plain C: pointers, sizes, ints and doubles only.  Host arrays use the
 * reference layout: AoS float[4*N], (x,y,z,w) per particle (MDSystem.h:72-74).
 * No function here has a CPU fallback: without a CUDA device every call that
 * needs one returns LJMD_ERR_CUDA.
 */
#ifndef LJMD_H
#define LJMD_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
#define LJMD_OK 0
#define LJMD_ERR_CUDA 1     /* a CUDA runtime call failed / no device          */
#define LJMD_ERR_ARG 2      /* invalid argument                                */
#define LJMD_ERR_NCCL 3     /* a NCCL call failed / built without NCCL         */
#define LJMD_ERR_DOMAIN 4   /* input outside the documented domain             */

/* Last error message of the calling thread ("" if none). */
const char* ljmd_last_error(void);

/* Boundary conditions: MDSystem.h:24-25. */
#define LJMD_BC_PERIODIC 0
#define LJMD_BC_HARDWALL 1
#define LJMD_BC_NONE 2

#define LJMD_RDF_BINS 256   /* MDSystem.cpp:97 */

/* Indices into the scalar block returned by ljmd_get_scalars
 * (public members of MDSystem, MDSystem.h:63-66,80,83,98-99). */
enum {
  LJMD_S_U = 0, LJMD_S_T, LJMD_S_K, LJMD_S_V, LJMD_S_P,
  LJMD_S_PVIRIAL,      /* virial part of P before CalculateParameters (MDSystem.cpp:307) */
  LJMD_S_TIME, LJMD_S_L,
  LJMD_S_AV_U_TOT, LJMD_S_AV_T_TOT, LJMD_S_AV_P_TOT, LJMD_S_AV_ITERS,
  LJMD_S_CHI,          /* last TVN rescale factor (MDSystem.cpp:499)             */
  LJMD_S_TKIN_TRIAL,   /* last KineticTemperature(t_Vel) (MDSystem.cpp:498)      */
  LJMD_S_COUNT = 16
};

typedef struct ljmd_system ljmd_system;

/* ------------------------------------------------- A. device-resident API */

/* Number of visible CUDA devices (0 when none / no driver). */
int ljmd_device_count(void);

/*
 * Create a system of N particles at density rho in a cubic box of edge
 * L = (N/rho)^(1/3) (MDSystem.cpp:70).  N, rho, T0, canonical, bc as in
 * MDSystemConfiguration (MDSystem.h:10-41).  rdf_dr2: bin width in r^2
 * (MDSystem.cpp:93-95; ljmd_rdf_dr2(N) restates that rule).
 * Replaces: MDSystem::Reinitialize/ReallocateMemory device part (MDSystem.cpp:116-144).
 */
int ljmd_create(ljmd_system** out, int N, double rho, double T0, int canonical, int bc,
                float rdf_dr2, int device);

/*
 * Same on `ndev` GPUs of one node driven by the ONE calling thread ("create(..., device list)").
 * The i-particles are split in contiguous shards over devices[0..ndev); every call on the returned handle is
 * forwarded to all devices and returns when the slowest is done.  No launcher, no IPC and no NCCL: the devices
 * map each other's memory (cudaDeviceEnablePeerAccess) and the per-step exchange — positions after the drift,
 * Newton-3 reaction forces, energy / virial / kinetic sums — runs inside the library's own kernels over NVLink.
 * Host arrays keep their full length: every device reads and writes its own shard of them.  ndev == 1 is
 * ljmd_create.  Fails when the devices cannot map each other's memory or N is too small to give every device a
 * 512-particle block.  New functionality: the reference is single-device.
 * Environment LJMD_SHARE_DEVICES=1 allows a device to be listed more than once (several ranks on one GPU, each on
 * its own stream): a diagnostic mode that runs the whole sharded data path on a box with fewer GPUs than ranks —
 * the ranks wait for each other inside kernels, so set CUDA_DEVICE_MAX_CONNECTIONS >= the rank count (one hardware
 * queue per rank's stream) and CUDA_MODULE_LOADING=EAGER (a lazily loaded kernel's first launch would synchronise
 * the context behind a peer's spinning barrier) before CUDA starts; without the latter the call fails with a message.
 */
int ljmd_create_multi(ljmd_system** out, int N, double rho, double T0, int canonical, int bc,
                      float rdf_dr2, const int* devices, int ndev);

/*
 * Same, as rank `rank` of `world` cooperating processes (one GPU each).  The
 * i-particles are split in contiguous shards; every rank holds all positions.
 * nccl_unique_id: the 128-byte ncclUniqueId produced by ljmd_nccl_unique_id on
 * rank 0 and shipped to the others by the caller (torch.distributed, MPI, ...).
 * New functionality: the reference is single-device.
 */
int ljmd_create_distributed(ljmd_system** out, int N, double rho, double T0, int canonical, int bc,
                            float rdf_dr2, int device, int rank, int world,
                            const void* nccl_unique_id);
int ljmd_nccl_unique_id(void* out128);

/*
 * Intra-node fabric (optional, after ljmd_create_distributed on every rank): each rank exports the 64-byte
 * CUDA-IPC handle of its window (positions, fixed-point records, reaction sums, reduction slots, flags),
 * the caller ships all handles to all ranks, and ljmd_fabric_connect maps the peers' windows.  From then on
 * the per-step exchange runs over NVLink peer memory inside the library's own kernels: k_drift stores each
 * new position straight into every rank's window (the all-gather is fused into the integrator),
 * k_gather pulls the peers' reaction sums for its own particles, and the scalar all-reduces are a
 * flag-and-slot barrier kernel with a fixed summation order.  NCCL stays in use for the rare read-out
 * collectives (ljmd_get_state, histograms) and as the transport when connect is never called or fails.
 * handles: world x 64 bytes in rank order.  All ranks must connect (or none).
 */
int ljmd_fabric_export(ljmd_system* s, void* out64);
int ljmd_fabric_connect(ljmd_system* s, const void* handles);

int ljmd_destroy(ljmd_system* s);

/* MDSystem.cpp:93-95. */
float ljmd_rdf_dr2(int N);

/* Live switches the callers flip between steps (MDSystem.h:160-163,
 * semiGCEfluctuations.cpp:58,66). */
int ljmd_set_canonical(ljmd_system* s, int canonical);
int ljmd_set_boundary(ljmd_system* s, int bc);
int ljmd_set_T0(ljmd_system* s, double T0);

/*
 * Upload a snapshot (host AoS float[4N] each; all ranks pass the full arrays)
 * and evaluate forces, V, virial, K, T, P, U on it; zero t and the av_*
 * accumulators.  Equivalent to writing h_Pos/h_Vel and calling
 * CalculateForces(); CalculateParameters(); resetAveraging()
 * (MDSystem.cpp:232-235,325-359,701-705) — the tail of Reinitialize (:99-113).
 * Coordinates may lie outside [0,L] (the minimum image handles any offset).
 */
int ljmd_set_state(ljmd_system* s, const float* pos4, const float* vel4);

/*
 * MDSystem::SampleInitialConditions on the device (MDSystem.cpp:147-181), then the evaluation ljmd_set_state does:
 * the reference's simple-cubic start lattice (bit-identical positions, w = L/150), Gaussian velocity components
 * N(0, T0) from a counter-based Philox4x32-10 generator keyed by `seed` and counted by the global particle index,
 * total momentum removed (CorrectTotalMomentum, :183-216) and velocities rescaled to T0 exactly
 * (RenormalizeVelocities(true), :375-389).  The same seed gives the same state bit for bit on any number of GPUs
 * (the two global sums are integer).  The reference's own generator is time-seeded, so there is nothing to be
 * bit-compatible with; the distribution is the same (Maxwell speeds, isotropic directions).
 */
int ljmd_init_state(ljmd_system* s, unsigned long long seed);

/* Plain upload of host arrays (either may be NULL) without any evaluation: what the reference does when
 * a caller edits h_Pos / h_Vel in place (copyArrayToDevice, MDSystem.cu:212-216). */
int ljmd_upload(ljmd_system* s, const float* pos4, const float* vel4);

/* Upload only velocities (after a host-side rescale, MDSystem.cpp:375-404) and
 * recompute K, T, P, U without touching the av_* accumulators or forces. */
int ljmd_set_velocities(ljmd_system* s, const float* vel4);

/* Download the current state into host AoS float[4N] arrays (any may be NULL).
 * pos4.w keeps the uploaded w (L/150, MDSystem.cpp:168); force4.w carries the
 * potential column of the reference GPU path (MDSystem.cu:52,135): sum over
 * particles of force4.w * 2 = V, which is all MDSystem.cpp:340-346 uses.  With the
 * ordered kernel the entry is the particle's own sum_j (r^-12 - r^-6); the
 * Newton-3 kernel books each unordered pair (twice) on the particle that held
 * it as "i", so only the sum is comparable.
 * Replaces copyArrayFromDevice (MDSystem.cu:199-210). */
int ljmd_get_state(ljmd_system* s, float* pos4, float* vel4, float* force4);

/*
 * nsteps x MDSystem::Integrate(dt) (MDSystem.cpp:438-583): drift, force
 * evaluation, EVN half-kick or TVN chi-rescale, boundary conditions,
 * CalculateParameters, t += dt.  Everything stays on the device; returns after
 * the last step completed.  Inside the batch the kernel finishing step k also
 * performs the drift of step k+1 (kick-drift-wrap fusion), and on one GPU the
 * steady-state steps of a long batch are replayed from a captured CUDA graph
 * (LJMD_GRAPH=0 disables it); results are bit-identical to nsteps single calls.
 * rdf_every > 0: the RDF histogram is rebuilt on
 * every rdf_every-th step of this call (and accumulated, see ljmd_get_rdf_accum);
 * 0: no RDF work (ljmd_get_rdf evaluates it lazily when asked).
 */
int ljmd_step(ljmd_system* s, double dt, int nsteps, int rdf_every);

/*
 * The drop-in single step with HOST buffers: upload pos/vel, one Integrate(dt),
 * download pos/vel (and forces when force4 != NULL).  This is what
 * MDSystem::Integrate costs a caller that reads h_Pos/h_Vel after every step.
 * Sharded systems: a device moves ITS shard both ways — up from the caller's
 * full-length arrays and back into the same places.  Behind ljmd_create_multi the
 * devices share the arrays, so the caller sees the whole state; with one process
 * per GPU (ljmd_create_distributed) each process gets its own shard refreshed
 * (33.5 MB instead of 268 MB of D2H per step at N = 1M on 8 GPUs) and fetches
 * the rest with ljmd_get_state when it wants it.
 */
int ljmd_integrate_host(ljmd_system* s, double dt, float* pos4, float* vel4, float* force4);

/* Re-evaluate forces/V/virial (and parameters) at the current positions:
 * MDSystem::CalculateForces + CalculateParameters without the av_* update. */
int ljmd_compute_forces(ljmd_system* s, int with_rdf);

/* out[LJMD_S_COUNT] doubles, see the enum above. */
int ljmd_get_scalars(ljmd_system* s, double* out);

/*
 * Shear stress P_xy, the one observable the reference computes on its CPU path only (MDSystem.cpp:299,309,335,353;
 * its GPU path leaves the member stale): (2 * sum_{i, j != i} (-r_x,ij f_y,ij / 4) + sum_i (-v_x v_y)) / (N / rho) for
 * the positions of the latest force evaluation and the current velocities.  Evaluated ON DEMAND by a separate
 * all-pairs pass (about the cost of one ordered force evaluation), so the step's hot loop does not pay for an
 * observable nothing in the reference reads.
 */
int ljmd_get_pshear(ljmd_system* s, double* pshear);

/*
 * Device-resident state of a single-device handle: float4 arrays of N entries (pos.w = L/150 as in h_Pos) and the
 * CUDA stream (cudaStream_t) the library orders its work on.  For renderers and downstream CUDA code that would
 * otherwise pay a D2H copy per frame (the reference's registerGLBufferObject hooks are stubs, MDSystem.cu:199-226).
 * Any pointer may be NULL.  Read-only for the caller; valid until ljmd_destroy.
 */
int ljmd_device_arrays(ljmd_system* s, const void** pos4, const void** vel4, const void** force4, void** stream);

/* MDSystem::resetAveraging (MDSystem.cpp:701-705). */
int ljmd_reset_averaging(ljmd_system* s);

/*
 * RDF histogram of the most recent force evaluation, CPU-path semantics
 * (MDSystem.cpp:279-285: bins in r^2, width rdf_dr2, pairs beyond bin 255
 * dropped, every unordered pair counted twice).  Bit-exact with that path.
 * Computed lazily from the saved evaluation positions if the last step did not
 * build it.  out256: int32 as NdNdr2 (MDSystem.h:90).
 */
int ljmd_get_rdf(ljmd_system* s, int* out256);

/* Sum of the histograms built by ljmd_step(..., rdf_every) since the last
 * reset, and how many evaluations went in.  reset != 0 clears afterwards. */
int ljmd_get_rdf_accum(ljmd_system* s, long long* out256, int* nsamples, int reset);

/*
 * Speed histogram counts (MDSystem.cpp:651-662,676-688):
 * bin = (int)(sqrt(vx^2+vy^2+vz^2) / step), dropped when >= nbins.  Bit-exact.
 */
int ljmd_velocity_histogram(ljmd_system* s, double step, int nbins, int* out);

/*
 * Cumulative sub-volume occupancies, counted on the device: what the fluctuation tasks'
 * GetNSubsystemBatch(syst, alpha_step, type) and GetNsubVzBatch(syst, vcut_max, alpha_step, type) compute from
 * h_Pos / h_Vel (src/tasks/run-fluctuations/include/run-fluctuations-aux.h:188-278).  type: 0/1/2 slab along
 * x/y/z, 3 cube about the centre; velocities 0/1/2 = |vx|,|vy|,|vz| < fraction * vcut_max.  out[k] = number of
 * particles in the sub-volume of fraction (k+1)*alpha_step; *nout = number of fractions (the reference's own
 * repeated-addition grid).  Bit-exact with those functions.
 */
int ljmd_subvolume_counts(ljmd_system* s, int type, double alpha_step, int* out, int cap, int* nout);
int ljmd_velocity_subvolume_counts(ljmd_system* s, int type, double vcut_max, double alpha_step, int* out, int cap,
                                   int* nout);

/* Kernel launches issued by this handle so far (for bench.py's gpu_launches). */
long long ljmd_launch_count(ljmd_system* s);

/* Milliseconds the last ljmd_step spent in its force kernels / in its steps, from CUDA events on the
 * handle's stream (0 when event timing is off).  With an L2 flush configured, total_ms is the sum of the
 * per-step intervals, i.e. it excludes the flush writes issued between steps. */
int ljmd_set_event_timing(ljmd_system* s, int on);
int ljmd_last_step_timing(ljmd_system* s, double* force_ms, double* total_ms, int* force_launches);

/* Benchmark hygiene: when bytes > 0, ljmd_step overwrites a scratch buffer of that size on the handle's
 * stream before every step, evicting the previous step's lines from the 126 MB L2. */
int ljmd_set_l2_flush(ljmd_system* s, long long bytes);

/* The same for k_gather, the dominant HBM-bound kernel of the step (sums the partial-force and reaction rows,
 * finishes the velocity update): total milliseconds and launches of the last call, and the algorithmic bytes
 * one launch moves. */
int ljmd_last_gather_timing(ljmd_system* s, double* gather_ms, int* launches, double* bytes_per_launch);

/* Sharded Newton-3 runs: the same for k_reduce_reaction, which sums this rank's reaction blocks into one record per
 * particle before the exchange (reads every block once: the HBM-bound kernel of the sharded step). */
int ljmd_last_reduce_timing(ljmd_system* s, double* reduce_ms, int* launches, double* bytes_per_launch);

/* FP32 CUDA-core throughput of `device` measured with a stream of independent packed FMAs (TFLOP/s, best of 5
 * launches, CUDA events): the measured denominator of the force kernel's roofline. */
int ljmd_fp32_peak_probe(int device, double* tflops);

/* Static facts for rooflines: out[0]=SM count, out[1]=i-tile size, out[2]=splits per i-tile,
 * out[3]=force CTAs per launch, out[4]=world size, out[5]=local particles, out[6]=1 if the Newton-3 kernel
 * (each unordered pair evaluated once) is in use, out[7]=j-records per shared-memory tile / work unit. */
int ljmd_get_launch_info(ljmd_system* s, int* out8);

/* Host-only helpers (no device needed), exported so the launch plan and the exact-RDF constants can
 * be checked without a GPU.
 * ljmd_image_threshold: smallest positive float d for which the reference's
 *   fast_round((float)(d / L)) (MDSystem.cpp:274,732-739) is >= k.
 * ljmd_plan: out[0..1] = i-shard [begin,end) of `rank` (whole 512-particle blocks), out[2] = i-tiles,
 *   out[3] = splits per i-tile, out[4] = force CTAs per launch, out[5] = i-particles per CTA,
 *   out[6] = 1 when the Newton-3 kernel is used, out[7] = partner blocks per i-tile, for `num_sms` SMs. */
float ljmd_image_threshold(double L, int k);
int ljmd_plan(int N, int rank, int world, int num_sms, int* out8);
/* Super-tile geometry of the Newton-3 kernel for the same arguments (csrc/ljmd_force_sym.cuh): out[0] = records per
 *   unit (0: the ordered kernel is used), out[1] = i-tiles per super-tile, out[2] = units per window, out[3] =
 *   windows per super-tile, out[4] = super-tiles of this rank, out[5] = launch-order shift of the windows,
 *   out[6] = global number of 512-particle blocks, out[7] = first global block of this rank. */
int ljmd_plan_newton3(int N, int rank, int world, int num_sms, int* out8);

/*
 * Observation trace: everything the fluctuation tasks read after EVERY step,
 * recorded on the device while ljmd_step runs and fetched in one copy.
 * Replaces the per-step host passes over h_Pos / h_Vel in
 * run-fluctuations.cpp:124-135 (CoordFlucsAverage / MomentumFlucsAverage ::
 * AddTimeStep -> GetNSubsystemBatch / GetNsubVzBatch,
 * run-fluctuations-aux.h:188-278), GetAvVel (:283-293) and the per-step reads
 * of U and P in run-isotherm.cpp:104-106 — which force a D2H of the whole state
 * per step — by one small kernel per step and one D2H per batch.
 *
 * kinds[c]: 0,1,2 slab in x,y,z; 3 centred cube (ljmd_subvolume_counts types);
 * 4,5,6 |vx|,|vy|,|vz| < vcut (ljmd_velocity_subvolume_counts types 0,1,2, which
 * need vcut_max[c]; the entry is ignored for kinds 0-3, vcut_max may be NULL
 * when there is no velocity counter).  At most 8 counters.  While a trace is
 * active ljmd_step appends one row per step and refuses to run past
 * capacity_steps unread rows; EVN batches run without the kick-drift fusion
 * (a row needs the end-of-step velocities; TVN has no first half-kick and keeps
 * it).  The row index lives on the device, so traced batches are replayed from
 * a CUDA graph like plain ones.
 */
int ljmd_trace_begin(ljmd_system* s, int ncounters, const int* kinds, const double* alpha_steps,
                     const double* vcut_max, int capacity_steps);
/* Number of counts per step: the bins of all counters, concatenated in order. */
int ljmd_trace_row_length(ljmd_system* s, int* counts_per_step);
/*
 * Fetch and clear the recorded rows (oldest first).  Any output may be NULL.
 *   scalars       [nsteps][LJMD_TRACE_SCALARS]: t, U, T, P, K, V, Pvirial, 0
 *   counts        [nsteps][row_length]: per counter the cumulative occupancies,
 *                 exactly what ljmd_subvolume_counts returns for that step
 *   mean_velocity [nsteps][3]: GetAvVel; summed in 2^-32 fixed point, so the
 *                 value does not depend on the launch shape or the GPU count
 */
#define LJMD_TRACE_SCALARS 8
int ljmd_trace_read(ljmd_system* s, int max_steps, int* nsteps, double* scalars, long long* counts,
                    double* mean_velocity);
int ljmd_trace_end(ljmd_system* s);

/* ------------------------------------------------- B. legacy seam ---------
 * The six symbols MDSystem.cpp calls (MDSystem.cpp:9-25; definitions replaced:
 * MDSystem.cu:167-173,181-184,199-216,230-291,294-297).  Same signatures, same
 * ownership (opaque device float* owned by the caller, host_RDF holds 256 ints).
 * Differences, all documented in INTEGRATION.md: RDF uses the CPU-path semantics
 * above; failures print to stderr and leave outputs zeroed instead of exit().
 */
void allocateArray(float** dest, int number);
void deleteArray(float* arr);
void copyArrayToDevice(float* device, const float* host, int numBodies);
void copyArrayFromDevice(float* host, const float* device, unsigned int pbo, int numBodies);
void calculateNForces(float* Pos, float* Force, float* host_pressure, int numBodies, float host_L,
                      int Lperiodic, int* host_RDF, float host_dr2, int p, int q);
void threadExit(void);
/* Declared by the reference host layer but never called (MDSystem.cpp:13,15,21-23). */
void allocateNBodyArrays(float* vel[2], int numBodies);
void deleteNBodyArrays(float* vel[2]);
void registerGLBufferObject(unsigned int pbo);
void unregisterGLBufferObject(unsigned int pbo);
void threadSync(void);

#ifdef __cplusplus
}
#endif
#endif /* LJMD_H */
