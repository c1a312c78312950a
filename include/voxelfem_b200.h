/* voxelfem_b200.h -- C ABI of the B200-native VoxelFEM hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The C++
 * host classes in voxelfem_b200/host/VoxelFEM.hh (TensorProductSimulator, MultigridSolver,
 * TopologyOptimizationProblem, ...) and the pyVoxelFEM / pyOptimizer modules in
 * voxelfem_b200/compat/ are thin veneers over these entry points; tests/ also call them
 * directly through ctypes.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * VoxelFEM reference tree).  All functions return 0 on success and a non-zero status on
 * failure; vf_last_error() then holds the message (the host wrappers rethrow it as the
 * std::runtime_error / std::logic_error the reference would have thrown).
 *
 * Conventions (identical to the reference):
 *   - grids are flattened row-major, axis 0 slowest;  nodes per dim = elements + 1 (Q1)
 *   - nodal vector fields cross this boundary as VField = (numNodes x N) column-major,
 *     i.e. component c of node n at data[c * numNodes + n]   (TensorProductSimulator.hh:180)
 *   - per-element scalar fields are flat arrays of numElements doubles
 *   - the build direction for fabrication masks is axis 1   (TensorProductSimulator.hh:1851)
 * Host pointers are pageable or pinned host memory unless the name ends in _dev.
 * There is no CPU fallback: every compute entry point runs CUDA kernels on the current
 * device and fails with an error status if no device is usable.
 */
#ifndef VOXELFEM_B200_H
#define VOXELFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vf_sim vf_sim; /* TensorProductSimulator<double,1,1[,1]>  (TensorProductSimulator.hh:170-2182) */
typedef struct vf_mg  vf_mg;  /* MultigridSolver<double,1,1[,1]>         (MultigridSolver.hh:23-1176)         */
typedef struct vf_top vf_top; /* TopologyOptimizationProblem + MultigridComplianceObjective + FilterChain +
                                 TotalVolumeConstraint + OCOptimizer state (TopologyOptimizationProblem.hh,
                                 TopologyOptimizationObjective.hh:72-105, OptimalityCriterion.hh:38-149)   */
typedef struct vf_lbl vf_lbl; /* LayerByLayerEvaluator (LayerByLayer.hh:25-309) */
typedef struct vf_gtop vf_gtop; /* TopologyOptimizationProblem + MultigridComplianceObjective + TotalVolumeConstraint + OCOptimizer on a slab group (no reference equivalent: the reference is single-address-space) */
typedef struct vf_group vf_group; /* a MultigridSolver partitioned into slabs along axis 0, one part per GPU (no reference equivalent: the reference is single-address-space) */

#define VF_LAW_SIMP 0
#define VF_LAW_RAMP 1
#define VF_FILTER_SMOOTH 0   /* SmoothingFilter  (TopologyOptimizationFilter.hh:285-400) */
#define VF_FILTER_PROJECT 1  /* ProjectionFilter (TopologyOptimizationFilter.hh:189-245) */
#define VF_FILTER_UPSAMPLE 2        /* UpsampleFilter      (:418-523); spec {2, factor, 0, 0} */
#define VF_FILTER_VERTEX_TO_CELL 3  /* VertexToCellFilter  (:528-598); spec {3, 0, 0, 0} */
#define VF_FILTER_LANGELAAR 4       /* LangelaarFilter     (:601-712); spec {4, 0, 0, 0} */
#define VF_FILTER_PYTHON 5          /* PythonFilter        (:247-275); spec {5, 0, 0, 0} + vf_top_set_python_filter */
#define VF_SMOOTH_CONST 0
#define VF_SMOOTH_LINEAR 1

/* ---- library / device ------------------------------------------------------------ */
const char *vf_last_error(void);
int vf_device_count(int *count);
int vf_set_device(int device);
int vf_version(void);
/* Number of kernels this library has launched since load / since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t vf_kernel_launch_count(void);
void vf_reset_kernel_launch_count(void);
/* Measured FP64 fused-multiply-add throughput of the current device in 1e12 DFMA/s (register-resident independent chains on
 * every SM): the denominator of the FP64-pipe roofline bench.py reports next to the HBM one ("FP64 peak to be
 * measured by the builder").  No reference equivalent. */
int vf_measure_fp64_peak(double *tera_dfma_per_s);

/* ---- section timers / tracing ----------------------------------------------------
 * MeshFEM's global benchmark timer (3rdParty/MeshFEM/src/lib/MeshFEM/GlobalBenchmark.hh, Timer.hh) as bound by
 * 3rdParty/MeshFEM/src/python_bindings/benchmark.cc:9-13.  Sections nest ("outer:inner"); the library opens the sections the
 * reference opens around the same steps (CG Iterations, Preamble, V Cycle <l>, updateStiffnessMatrices, OC step, Bisection,
 * setVars, Build load, Update load, Construct initial guess, Compute compliance, Compute gradient) and emits NVTX ranges of
 * the same names.  Timing is off by default (a section drains the device when it closes): vf_benchmark_enable(1) or
 * VF_BENCHMARK=1 turns it on. */
void vf_benchmark_enable(int on);
int vf_benchmark_enabled(void);
void vf_benchmark_reset(void);
void vf_benchmark_start_timer_section(const char *name);
void vf_benchmark_stop_timer_section(const char *name);
void vf_benchmark_start_timer(const char *name);
void vf_benchmark_stop_timer(const char *name);
void vf_benchmark_add_message(const char *msg);
/* one line per section / timer: "path<TAB>seconds<TAB>(invocations)"; returns the length needed (excluding the 0) */
size_t vf_benchmark_report(int include_messages, char *buf, size_t capacity);

/* ---- TensorProductSimulator ------------------------------------------------------ */
/* ctor, TensorProductSimulator.hh:209-279.  dim = 2 or 3. */
int vf_sim_create(int dim, const int64_t *ne, const double *domain_min, const double *domain_max, vf_sim **out);
int vf_sim_destroy(vf_sim *s);
int64_t vf_sim_num_nodes(const vf_sim *s);
int64_t vf_sim_num_elements(const vf_sim *s);
/* setETensor (:343-347) with the flattened tensor D (3x3 in 2D, 6x6 in 3D, row-major, MeshFEM
 * Voigt order xx,yy[,zz,yz,xz],xy); recomputes K0 by 2-point Gauss quadrature (:67-80, :2078-2086). */
int vf_sim_set_elasticity_tensor(vf_sim *s, const double *D);
/* ElasticityTensor::setIsotropic (MeshFEM ElasticityTensor.hh:100-115; 2D = plane stress). */
int vf_sim_set_isotropic(vf_sim *s, double young, double poisson);
/* fullDensityElementStiffnessMatrix (:1077): (N*2^N)^2 doubles, row-major. */
int vf_sim_get_K0(const vf_sim *s, double *out);
/* setInterpolationLaw / setE_0 / setE_min / setSIMPExponent / setRAMPFactor (:1719-1723). */
int vf_sim_set_interpolation(vf_sim *s, int law, double E_0, double E_min, double gamma, double q);
int vf_sim_set_gravity(vf_sim *s, const double *g);                      /* setGravity (:1775) */
int vf_sim_set_densities(vf_sim *s, const double *rho);                  /* setDensities (:785-790) */
int vf_sim_set_uniform_density(vf_sim *s, double rho);                   /* setUniformDensities (:715-720) */
int vf_sim_get_densities(const vf_sim *s, double *rho);                  /* getDensities (:780-783) */
int vf_sim_get_young_moduli(const vf_sim *s, double *E);                 /* getYoungModulusScaleFactor (:969) */
/* applyDisplacementsAndLoads (:600-652) for axis-aligned box regions given in absolute coordinates
 * (3 doubles per corner / value, unused trailing entries ignored).  kind[r]: 0 = dirichlet with
 * component bit-mask cmask[r] (bit c = component c), 1 = force (total force split evenly over the nodes
 * in the box, :629-630). */
int vf_sim_apply_bc_regions(vf_sim *s, int nregions, const int32_t *kind, const int32_t *cmask,
                            const double *values, const double *box_min, const double *box_max);
/* addDirichletCondition (:660-671). */
int vf_sim_add_dirichlet_box(vf_sim *s, const double *u, const double *box_min, const double *box_max, int cmask);
/* applySymmetryConditions (:2042-2053): axes / minMaxFace are bit-masks over the N axes. */
int vf_sim_apply_symmetry_conditions(vf_sim *s, int axes_mask, int max_face_mask);
int vf_sim_get_dirichlet_mask(const vf_sim *s, uint8_t *mask_per_node);  /* getDirichletMask (:675-682), bit c = component c */
int64_t vf_sim_num_force_nodes(const vf_sim *s);
int64_t vf_sim_num_dirichlet_nodes(const vf_sim *s);
/* m_dirichletNodes / m_dirichletComponents / m_dirichletNodeDisplacements (getDirichletVarsAndValues, TensorProductSimulator.hh:1790-1835;
 * getBCIndicatorField :697-712): nodes[n], masks[n] (bit c = component c constrained), values[n * N] */
int vf_sim_get_dirichlet_conditions(const vf_sim *s, int64_t *nodes, uint8_t *masks, double *values);
/* m_forceNodes / m_forceNodeForces (getForceMask :685-695): nodes[n], forces[n * N], n = vf_sim_num_force_nodes() */
int vf_sim_get_force_nodes(const vf_sim *s, int64_t *nodes, double *forces);
int64_t vf_sim_num_nonzero_dirichlet_values(const vf_sim *s);  /* prescribed non-zero displacements (getIntermediateFabricationShape validates them, :1893-1895) */
int vf_sim_build_load_vector(vf_sim *s, double *f);                      /* buildLoadVector (:1269-1288) */
/* applyK<ZeroInit,Negate> (:1410-1438 -> TPSStencils.hh:231-396, 431-728).
 * zero_init=1: out = K u; zero_init=0: out +=/-= K u (out is read). */
int vf_sim_apply_K(vf_sim *s, const double *u, double *out, int zero_init, int negate);
int vf_sim_set_mask_layer(vf_sim *s, int64_t layer);                     /* setFabricationMaskHeightByLayer (:327-329) */
int vf_sim_get_mask_info(const vf_sim *s, int64_t *first_masked_elem_layer, int64_t *first_detached_node_layer, double *height);
/* complianceGradientFlattened (:1051-1055) / accumulateComplianceGradient (:1008-1040). */
int vf_sim_compliance_gradient(vf_sim *s, const double *u, double *g, int accumulate);
int vf_sim_element_energy_density(vf_sim *s, const double *u, double *out); /* elementEnergyDensity (:1057-1073) */
/* TPS::solve (:1198-1230): direct solve with a dense GPU Cholesky; intended for coarse grids only
 * (fails for more than VF_MAX_DIRECT_DOFS free variables). */
#define VF_MAX_DIRECT_DOFS 20000
int vf_sim_solve(vf_sim *s, const double *f, double *u);

/* ---- MultigridSolver --------------------------------------------------------------- */
/* ctor (MultigridSolver.hh:35-121): builds the hierarchy, coarsens the Dirichlet conditions (:58-103)
 * and the 2^N coarsened full-density matrices (:116-120).  The simulator must outlive the solver and
 * is mutated by it (mask height), exactly as in the reference. */
int vf_mg_create(vf_sim *fine, int num_coarsening_levels, vf_mg **out);
int vf_mg_destroy(vf_mg *mg);
int vf_mg_num_levels(const vf_mg *mg);
int64_t vf_mg_level_num_nodes(const vf_mg *mg, int level);
int vf_mg_level_grid(const vf_mg *mg, int level, int64_t *ne);
int vf_mg_level_dirichlet_mask(const vf_mg *mg, int level, uint8_t *mask_per_node);
int vf_mg_get_coarsened_fine_K0(const vf_mg *mg, int fi, double *out);   /* coarsenedFineK0s (:1161) */
int vf_mg_update_stiffness_matrices(vf_mg *mg);                           /* updateStiffnessMatrices (:846-905), banded variant (:907-1017) */
int vf_mg_apply_K(vf_mg *mg, int level, const double *u, double *out);    /* applyK(l, u) (:464-504) */
int vf_mg_compute_residual(vf_mg *mg, int level, const double *u, const double *b, double *r); /* computeResidual (:527-541) */
int vf_mg_smooth(vf_mg *mg, int level, double *u, const double *b, int forward);  /* smoothingMulticoloredGS (:452-458) */
/* smoothingMulticoloredGS followed by computeResidual (:452-458, :527-541) as the V-cycle runs them on a stored-stencil level
 * (level >= 1, no detached layers): one sweep that also leaves r = b - K u of its final iterate (Dirichlet components zero).
 * Fails with an error status where the fused form is not available (level 0, fabrication mask active). */
int vf_mg_smooth_residual(vf_mg *mg, int level, double *u, const double *b, int forward, double *r);
int vf_mg_restrict(vf_mg *mg, int fine_level, const double *fine, double *coarse);            /* restriction (:216-262) */
int vf_mg_interpolate(vf_mg *mg, int fine_level, const double *coarse, double *fine, int accumulate); /* interpolation / accum_interpolation (:178-212) */
/* Assembled 3^N-point block stencil of a coarse level: [node][3^N][N][N] doubles (the reference's
 * blockK, TensorProductSimulator.hh:885-966, in dense-slot form).  level >= 1. */
int vf_mg_get_stencil(vf_mg *mg, int level, double *out);
int vf_mg_coarse_solve(vf_mg *mg, const double *f, double *x);            /* coarsest TPS::solve, (:622-624) */
/* solve (:546-573): numSteps V-cycles (first one a full-multigrid cycle if fmg). */
int vf_mg_solve(vf_mg *mg, const double *u, const double *f, int num_steps, int num_smoothing_steps,
                int stiffness_updated, int zero_dirichlet, int fmg, double *out);
/* preconditionedConjugateGradient (:1047-1152).  x is updated in place.  residual_norms (may be NULL)
 * receives ||r|| after every iteration (what the reference's it_callback would compute), at most max_iter
 * entries.  cb (may be NULL) is invoked after every iteration with (iteration, ||r||, user). */
typedef void (*vf_pcg_callback)(int iteration, double residual_norm, void *user);
int vf_mg_pcg(vf_mg *mg, double *x, const double *b, int max_iter, double tol, int mg_iterations,
              int mg_smoothing_iterations, int fmg, int dirichlet_already_satisfied,
              int *out_iterations, double *residual_norms, vf_pcg_callback cb, void *user);
/* Out-of-place form of vf_mg_pcg, as the reference's Python binding uses the solver (VoxelFEM.cc:174-186): u0 (initial guess)
 * and b are read, the solution is written to x_out. */
int vf_mg_pcg_io(vf_mg *mg, const double *u0, const double *b, double *x_out, int max_iter, double tol, int mg_iterations,
                 int mg_smoothing_iterations, int fmg, int dirichlet_already_satisfied,
                 int *out_iterations, double *residual_norms, vf_pcg_callback cb, void *user);
int vf_mg_get_pcg_residual(vf_mg *mg, double *r);                         /* pcgResidual (:1156) */
/* Current iterate x of the running PCG: valid inside cb (the reference's it_callback receives (it, x, r),
 * MultigridSolver.hh:1043-1045, 1146-1147) and until the buffers passed to the last solve are released. */
int vf_mg_get_pcg_iterate(vf_mg *mg, double *x);
int vf_mg_set_symmetric_gauss_seidel(vf_mg *mg, int symmetric);           /* setSymmetricGaussSeidel (:123-125) */
/* on != 0: every preconditionedConjugateGradient call rebuilds the coarse hierarchy in its first iteration, as the reference does
 * (stiffnessMatricesUpdated = false, MultigridSolver.hh:1104-1107); default: only when the moduli or the mask changed since the last
 * build.  The host-buffer entry point vf_mg_pcg_io runs that rebuild while its input copies are in flight. */
int vf_mg_set_rebuild_every_solve(vf_mg *mg, int on);
int vf_mg_set_mask_layer(vf_mg *mg, int64_t fine_layer);                  /* setFabricationMaskHeightByLayer (:1022-1028) */
int vf_mg_decrement_mask(vf_mg *mg, int fine_layer_increment);            /* decrementFabricationMaskHeightByLayer (:1030-1036) */
int vf_mg_debug_get(vf_mg *mg, int which /*0 x, 1 b, 2 r*/, int level, double *out); /* debug_get_x/b (:1158-1159) */
int vf_mg_debug_multicolor_visit(vf_mg *mg, int32_t *order_per_node);     /* debugMulticolorVisit (:444-450) */

/* Device-resident variants used by the optimization layer and by bench.py's HBM-resident timing:
 * x_dev / b_dev are device pointers to VFields allocated with vf_dev_alloc. */
int vf_dev_alloc(size_t num_doubles, double **out_dev);
int vf_dev_free(double *dev);
int vf_dev_upload(double *dev, const double *host, size_t num_doubles);
int vf_dev_download(double *host, const double *dev, size_t num_doubles);
int vf_dev_memset_zero(double *dev, size_t num_doubles);
int vf_mg_pcg_dev(vf_mg *mg, double *x_dev, const double *b_dev, int max_iter, double tol, int mg_iterations,
                  int mg_smoothing_iterations, int fmg, int dirichlet_already_satisfied,
                  int *out_iterations, double *residual_norms, vf_pcg_callback cb, void *user);
int vf_sim_build_load_vector_dev(vf_sim *s, double *f_dev);
/* The CUDA stream all kernels of this solver are launched on (as a cudaStream_t). */
void *vf_mg_stream(vf_mg *mg);
int vf_mg_synchronize(vf_mg *mg);

/* Per-kernel device timing (CUDA events on the solver's stream), for the roofline report.
 * Categories: see vf_prof_name().  Enabling adds two event records per launch. */
int vf_prof_enable(vf_mg *mg, int enable);
int vf_prof_reset(vf_mg *mg);
int vf_prof_num_categories(void);
const char *vf_prof_name(int category);
int vf_prof_get(vf_mg *mg, int category, int64_t *launches, double *total_ms, double *units);
/* Device time (ms) of one repetition of a single multigrid operation, averaged over `reps` back-to-back
 * repetitions on the level's own fields.  op: 0 smoothing sweep (smoothingMulticoloredGS, :452-458), 1 computeResidual,
 * 2 applyK, 3 restriction, 4 accum_interpolation, 5 coarsest solve, 6 vcycle(level) (:617-658), 7 fullMultigrid(0) (:587-609),
 * 8 forward smoothing sweep that also leaves computeResidual's result (what vcycle runs on a stored-stencil level, level >= 1). */
int vf_mg_time_op(vf_mg *mg, int op, int level, int reps, int num_smoothing_steps, double *ms_per_rep);

/* ---- Slab-partitioned solver (multi-GPU) -------------------------------------------------
 * The grid is cut into slabs of element layers [slab_begin, slab_end) along axis 0 (the slowest axis: a slab and its halo
 * planes are contiguous in every field).  Each part is an ordinary simulator / solver over the slab's WINDOW: the node planes
 * [max(slab_begin - 1, 0), min(slab_end + 1, ne[0])] -- the slab plus one ghost plane per neighbour; per-element and nodal
 * arrays passed to a part (densities, u, f) are the window's slice of the global arrays.  BC regions are matched against the
 * global grid (a force is split over all nodes of the grid inside its box).  Levels < first_replicated_level are windowed
 * per part, the others are held by every part for the whole grid; slab boundaries must be multiples of
 * 2^first_replicated_level.  One process per GPU joins an NCCL group (halo planes by ncclSend/ncclRecv, PCG scalars and
 * the replicated coarse data by ncclAllReduce); a local group drives several parts from one process on one device. */
int vf_sim_create_slab(int dim, const int64_t *ne_global, const double *domain_min, const double *domain_max,
                       int64_t slab_begin, int64_t slab_end, vf_sim *share_stream_with /* may be NULL */, vf_sim **out);
/* node planes stored [plane_lo, plane_hi] and owned [own_lo, own_hi] (global indices along axis 0) */
int vf_sim_window(const vf_sim *s, int64_t *plane_lo, int64_t *plane_hi, int64_t *own_lo, int64_t *own_hi);
int vf_mg_create_slab(vf_sim *fine, int num_coarsening_levels, int first_replicated_level, vf_mg **out);
int vf_group_create_local(int nparts, vf_mg **parts, vf_group **out);
int vf_nccl_unique_id(void *out128);  /* rank 0: ncclGetUniqueId; the caller broadcasts the 128 bytes (torch.distributed) */
int vf_group_create_nccl(vf_mg *part, int rank, int world, const void *unique_id128, vf_group **out);
int vf_group_destroy(vf_group *g);
/* preconditionedConjugateGradient (MultigridSolver.hh:1047-1152) over the whole grid; x_dev / b_dev: one device VField of the
 * part's window per local part.  Every part returns the same iteration count and residual history. */
int vf_group_pcg_dev(vf_group *g, double *const *x_dev, const double *const *b_dev, int max_iter, double tol, int mg_iterations,
                     int mg_smoothing_iterations, int fmg, int dirichlet_already_satisfied,
                     int *out_iterations, double *residual_norms, vf_pcg_callback cb, void *user);

/* ---- Filters (stand-alone) -------------------------------------------------------- */
int vf_filter_smooth(int dim, const int64_t *sizes, int radius, int type, const double *in, double *out); /* SmoothingFilter::apply == backprop (:297-310) */
int vf_filter_project(int64_t n, double beta, const double *in, double *out);                                 /* ProjectionFilter::apply (:199-210) */
int vf_filter_project_backprop(int64_t n, double beta, const double *in, const double *vars, double *out);  /* ProjectionFilter::backprop (:212-225) */
/* UpsampleFilter::apply / backprop (:418-523): a vertex grid of coarse_sizes <-> (coarse_sizes - 1) * factor + 1 */
int vf_filter_upsample(int dim, const int64_t *coarse_sizes, int factor, const double *in, double *out);
int vf_filter_upsample_backprop(int dim, const int64_t *coarse_sizes, int factor, const double *d_dout, double *d_din);
/* VertexToCellFilter::apply / backprop (:528-598): vertex_sizes <-> vertex_sizes - 1 */
int vf_filter_vertex_to_cell(int dim, const int64_t *vertex_sizes, const double *in, double *out);
int vf_filter_vertex_to_cell_backprop(int dim, const int64_t *vertex_sizes, const double *d_dout, double *d_din);
/* LangelaarFilter::apply / backprop (:601-712).  `out` is IN/OUT (in 3D a voxel's support, NDVector.hh:211-229, holds the voxel
 * itself: the previous content of the output array is read); smax (may be NULL) receives m_cachedSmax.  backprop takes the
 * filter's input (vars), its last output (filtered = m_cachedFiltered) and that smax. */
int vf_filter_langelaar(int dim, const int64_t *sizes, const double *in, double *out, double *smax);
int vf_filter_langelaar_backprop(int dim, const int64_t *sizes, const double *d_dout, const double *vars, const double *filtered, const double *smax, double *d_din);

/* ---- Device-pointer building blocks (slab-partitioned topology optimization) -------------
 * Same kernels as the filters / OC update / sensitivities above, on arrays that already live in HBM and on the
 * simulator's stream: the host side (voxelfem_b200/capi.py: SlabProblem) exchanges the filter halos between slabs
 * and all-reduces the scalars between these calls.  sizes / n describe the array passed in (e.g. a slab plus its halo). */
int vf_dev_filter_smooth(vf_sim *s, int dim, const int64_t *sizes, int radius, int type, const double *in_dev, double *out_dev);
int vf_dev_filter_project(vf_sim *s, int64_t n, double beta, const double *in_dev, double *out_dev);
int vf_dev_filter_project_backprop(vf_sim *s, int64_t n, double beta, const double *g_dev, const double *vars_dev, double *out_dev);
/* OCOptimizer update rule (OptimalityCriterion.hh:64-83) for one multiplier value. */
int vf_dev_oc_update(vf_sim *s, int64_t n, const double *x0_dev, const double *dJ_dev, const double *dc_dev, double lambda, double m, double p, double *out_dev);
int vf_dev_sum(vf_sim *s, int64_t n, const double *x_dev, double *result_host);
int vf_sim_set_densities_dev(vf_sim *s, const double *rho_dev);                 /* setDensities (:785-790), window-sized */
int vf_sim_compliance_gradient_dev(vf_sim *s, const double *u_dev, double *g_dev, int accumulate); /* (:1008-1040) */
void *vf_sim_stream(vf_sim *s);
int vf_sim_synchronize(vf_sim *s);

/* ---- Topology optimization problem ------------------------------------------------ */
/* filter_spec: 4 doubles per filter (kind, radius, smoothing type, beta). */
/* PythonFilter callbacks (:253-254): apply(in, out), backprop(d_dout, vars, d_din); return non-zero to abort */
typedef int (*vf_filter_apply_cb)(const double *in, int64_t n_in, double *out, int64_t n_out, void *user);
typedef int (*vf_filter_backprop_cb)(const double *d_dout, int64_t n_out, const double *vars, int64_t n_in, double *d_din, void *user);
int vf_top_create(vf_mg *mg, int num_filters, const double *filter_spec, double volume_fraction, vf_top **out);
int64_t vf_top_num_vars(const vf_top *t);                       /* FilterChain::numVars (:134): design variables (differs from numElements with Upsample / VertexToCell filters) */
int64_t vf_top_num_physical_vars(const vf_top *t);              /* numPhysicalVars (:138) */
int vf_top_get_grid_dims(const vf_top *t, int physical, int64_t *dims); /* gridDims / physicalGridDims (:136, 139) */
int vf_top_set_python_filter(vf_top *t, int filter_index, vf_filter_apply_cb apply_cb, vf_filter_backprop_cb backprop_cb, void *user);
/* MultigridComplianceObjective::residual_cb (TopologyOptimizationObjective.hh:93, 104): called after every PCG iteration of the
 * solves setVars / the OC step run, with the iteration number and residual norm; inside it vf_mg_get_pcg_residual gives r. */
int vf_top_set_residual_callback(vf_top *t, vf_pcg_callback cb, void *user);
int vf_top_destroy(vf_top *t);
/* MultigridComplianceObjective attributes (TopologyOptimizationObjective.hh:99-103). */
int vf_top_set_solver(vf_top *t, int cg_iter, double tol, int mg_iterations, int mg_smoothing_iterations, int fmg, int zero_init);
int vf_top_set_vars(vf_top *t, const double *x);                 /* setVars (TopologyOptimizationProblem.hh:41-50) */
int vf_top_get_vars(vf_top *t, int which /*0 design, 1 physical*/, double *out);
int vf_top_compliance(vf_top *t, double *out);                    /* ComplianceObjective::compliance (:41-43) */
int vf_top_constraint(vf_top *t, double *out);                    /* TotalVolumeConstraint::evaluate (TopologyOptimizationConstraint.hh:30-32) */
int vf_top_objective_gradient(vf_top *t, double *g);              /* evaluateObjectiveGradient (:77-84) */
int vf_top_constraint_jacobian(vf_top *t, double *g);             /* evaluateConstraintsJacobian (:106-121), single row */
int vf_top_get_u(vf_top *t, double *u);
int vf_top_last_pcg_iterations(vf_top *t);
int vf_top_oc_step(vf_top *t, double m, double p, double ctol, int *num_constraint_evals); /* OCOptimizer::step (OptimalityCriterion.hh:51-134) */
/* Search half of OCOptimizer::step for problems whose virtual methods are overridden on the host (trampoline,
 * python_bindings/VoxelFEM.cc:58-66; step(inplace = false), OptimalityCriterion.hh:57-60): dJ (host, numVars; NULL = the
 * problem's own compliance gradient) is the caller's evaluateObjectiveGradientAndReturn(); the bracket/bisection (:95-129) runs
 * on the device against the problem's own filter chain and constraint (evaluateOCConstraintAtVars is not virtual); the stepped
 * variables come back in `stepped` (host) and are NOT set -- the caller invokes its own setVars (:133). */
int vf_top_oc_search(vf_top *t, const double *dJ, double m, double p, double ctol, double *stepped, int *num_constraint_evals);
int vf_top_get_lambda_bracket(vf_top *t, double *lo, double *hi);

/* ---- compliance topology optimization on a slab group (BASELINE.json configs[3]) ---------------------------------------
 * TopologyOptimizationProblem.hh:17-155 + OptimalityCriterion.hh:38-149 with the element arrays partitioned like the solver:
 * every part owns the design variables of its element layers; per evaluation of the filter chain the parts exchange
 * vf_group_top_halo_layers() element layers with their neighbours (device copies in a local group, ncclSend/Recv between NCCL
 * ranks), volume and compliance are all-reduced and every rank runs the same bracket / bisection.  Filter spec as for
 * vf_top_create (Smoothing and Projection).  Whole-grid arrays cross the boundary: x / gradients are host arrays over the
 * GLOBAL element grid, identical on every rank. */
int vf_group_top_create(vf_group *g, int num_filters, const double *filter_spec, double volume_fraction, vf_gtop **out);
int vf_group_top_destroy(vf_gtop *t);
int64_t vf_group_top_halo_layers(const vf_gtop *t);
int vf_group_top_set_solver(vf_gtop *t, int cg_iter, double tol, int mg_iterations, int mg_smoothing_iterations, int fmg, int zero_init);
int vf_group_top_set_vars(vf_gtop *t, const double *x_global);                 /* setVars (:41-50) */
int vf_group_top_get_vars(vf_gtop *t, int which, double *out_global);          /* 0: getVars, 1: getDensities */
int vf_group_top_compliance(vf_gtop *t, double *out);                          /* evaluateObjective */
int vf_group_top_constraint(vf_gtop *t, double *out);                          /* evaluateConstraints()[0] */
int vf_group_top_objective_gradient(vf_gtop *t, double *g_global);             /* evaluateObjectiveGradient */
int vf_group_top_constraint_jacobian(vf_gtop *t, double *g_global);            /* evaluateConstraintsJacobian row 0 */
int vf_group_top_last_pcg_iterations(vf_gtop *t);
int vf_group_top_get_u(vf_gtop *t, int local_part, double *u_window);          /* displacement window of a local part, component-major */
int vf_group_top_oc_step(vf_gtop *t, double m, double p, double ctol, int *num_constraint_evals); /* OCOptimizer::step (OptimalityCriterion.hh:51-134) */

/* ---- Degree-2 (Q2) elements: TensorProductSimulator<double, 2, 2[, 2]> on the reference's generic element path ---------------
 * (the reference's python bindings instantiate degree 1 only, python_bindings/VoxelFEM.cc:303-308).
 * Node grid (2 ne + 1)^N; nodal fields component-major (c * numNodes + node) like everywhere in this ABI; K0 is (N 3^N)^2 with
 * entry index N * local_node + component, local nodes row-major over 3^N (TPSStencils.hh:139).                                   */
typedef struct vf_q2 vf_q2;
int vf_q2_create(int dim, const int64_t *num_elements, const double *domain_min, const double *domain_max, vf_q2 **out);   /* TensorProductSimulator.hh:209-279 */
int vf_q2_destroy(vf_q2 *s);
int64_t vf_q2_num_nodes(const vf_q2 *s);
int64_t vf_q2_num_elements(const vf_q2 *s);
int vf_q2_set_isotropic(vf_q2 *s, double young, double poisson);                       /* setETensor (:343-347) */
int vf_q2_set_elasticity_tensor(vf_q2 *s, const double *D);                            /* flattened 6x6 / 3x3, row-major */
int vf_q2_get_K0(const vf_q2 *s, double *out);                                         /* fullDensityElementStiffnessMatrix: Element_T::Stiffness (:67-80), Gauss degree 4 per axis */
int vf_q2_set_interpolation(vf_q2 *s, int law, double E_0, double E_min, double gamma, double q);   /* :2055-2102 */
int vf_q2_set_densities(vf_q2 *s, const double *rho);
int vf_q2_get_young_moduli(const vf_q2 *s, double *E);
int vf_q2_apply_K(vf_q2 *s, const double *u, double *out, int zero_init, int negate);  /* generic applyK<ZeroInit, Negate> (TPSStencils.hh:163-185) */
int vf_q2_element_energies(vf_q2 *s, const double *u, double *energy);                 /* u_e^T K0 u_e per element (elementEnergyDensity :1057-1073 up to the modulus) */
/* Jacobi-preconditioned CG on the Q2 operator with the flagged components (numNodes * N bytes, component-major) clamped to zero;
 * x: initial guess in, solution out.  Stands in for the reference's direct solve of Q2 systems (TPS::solve, :1198-1230): the
 * reference has no multigrid instantiation for degree 2 either. */
int vf_q2_pcg(vf_q2 *s, double *x, const double *b, const uint8_t *fixed, int max_iter, double tol, int *iters, double *rel_residual);

/* ---- Layer-by-layer evaluator ------------------------------------------------------ */
int vf_lbl_create(vf_mg *mg, vf_lbl **out);                       /* LayerByLayerEvaluator(lblSim) (LayerByLayer.hh:33-36) */
int vf_lbl_destroy(vf_lbl *l);
int vf_lbl_select_init_method(vf_lbl *l, const char *method);    /* selectInitMethod (:214-220): "zero", "constant", "fd", "N=k" */
typedef void (*vf_lbl_callback)(int64_t layer, double compliance, int pcg_iterations, void *user);
/* run (:223-296).  cb is the reference's lblCallback(l, compliance, grad_compliance, u) (:222, 277-279): the two arrays are not
 * pushed through the callback but fetched on demand from inside it with vf_lbl_get_layer_gradient / vf_lbl_get_layer_u (no copy
 * when the callback does not look at them).  pcg_cb is the it_callback handed to every layer's PCG (:265); inside it
 * vf_mg_get_pcg_iterate / vf_mg_get_pcg_residual give (x, r) as for vf_mg_pcg. */
int vf_lbl_run(vf_lbl *l, int zero_init, int64_t layer_increment, int max_iter, double tol, int mg_iterations,
               int mg_smoothing_iterations, int fmg, vf_lbl_callback cb, void *user, vf_pcg_callback pcg_cb, void *pcg_user);
int vf_lbl_get_layer_u(vf_lbl *l, double *u);                     /* u of the layer just solved, component-major (numNodes x N) */
int vf_lbl_get_layer_gradient(vf_lbl *l, double *g);              /* complianceGradientFlattened(u) of that layer (:278) */
int vf_lbl_objective(vf_lbl *l, double *out);                     /* objective (:299) */
int vf_lbl_gradient(vf_lbl *l, double *g);                        /* gradient (:300) */

/* ---- Layer-by-layer evaluator on a slab group (LayerByLayer.hh:25-309 with the grid partitioned along axis 0; the build direction is
 * axis 1, so every slab holds a piece of every layer).  Every rank calls the functions with the same arguments; the gradient
 * crosses the boundary as an array over the WHOLE element grid. */
typedef struct vf_glbl vf_glbl;
int vf_group_lbl_create(vf_group *g, vf_glbl **out);
int vf_group_lbl_destroy(vf_glbl *l);
int vf_group_lbl_select_init_method(vf_glbl *l, const char *method);            /* selectInitMethod (:214-220) */
int vf_group_lbl_run(vf_glbl *l, int zero_init, int64_t layer_increment, int max_iter, double tol, int mg_iterations,
                     int mg_smoothing_iterations, int fmg, vf_lbl_callback callback, void *user);   /* run (:223-296) */
int vf_group_lbl_objective(vf_glbl *l, double *out);                            /* objective (:299) */
int vf_group_lbl_gradient(vf_glbl *l, double *g_global);                        /* gradient (:300) */
int vf_group_lbl_num_layer_iterations(vf_glbl *l, int *iters);                  /* PCG iterations per simulated layer; returns their number */


/* ---- Method of Moving Asymptotes (pyOptimizer.MMA, python_bindings/Optimizer.cc:11-23) ----------------------
 * MMA(numVars, numConstr, xmin, xmax, f, df_dx) (MethodOfMovingAsymptotes.hh:34-52).  f writes the m + 1 values
 * (objective, constraints f_i(x) <= 0) and df_dx the (m + 1) x n row-major gradients; both return 0 on success.
 * With callbacks_take_device_pointers != 0 the callbacks receive / fill DEVICE arrays (x, df), so an objective that
 * lives on the GPU (vf_top / vf_lbl) never crosses the bus; otherwise they are host arrays as in the reference. */
#define VF_MMA_MAX_CONSTRAINTS 8
typedef struct vf_mma vf_mma;
typedef int (*vf_mma_f_callback)(const double *x, double *f_out, void *user);
typedef int (*vf_mma_df_callback)(const double *x, double *df_out, void *user);
int vf_mma_create(int64_t num_vars, int num_constr, const double *xmin, const double *xmax, vf_mma **out);
int vf_mma_destroy(vf_mma *mma);
int vf_mma_enable_gcmma(vf_mma *mma, int enable);                 /* enableGCMMA (:53) */
int vf_mma_set_initial_var(vf_mma *mma, const double *x);         /* setInitialVar (:54-56) */
int vf_mma_step(vf_mma *mma, vf_mma_f_callback f, vf_mma_df_callback df_dx, void *user, int callbacks_take_device_pointers); /* step (:63-133) */
int vf_mma_get_optimal_var(vf_mma *mma, double *x);               /* getOptimalVar (:134) */
int vf_mma_get_optimal_var_dev(vf_mma *mma, const double **x_dev);
int64_t vf_mma_newton_iterations(const vf_mma *mma);              /* interior-point Newton directions computed so far */

#ifdef __cplusplus
}
#endif
#endif /* VOXELFEM_B200_H */
