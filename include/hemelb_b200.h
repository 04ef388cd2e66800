/* hemelb_b200.h -- C ABI of the B200-native collide-and-stream engine.
 *
 * This is the drop-in boundary for HemeLB's lattice-Boltzmann hot path.  The reference has no
 * FFI: its seam is the C++ policy-template API (lb::streamer concept, Code/lb/concepts.h:90-103;
 * Traits<>, Code/Traits.h:17-39; geometry::FieldData, Code/geometry/FieldData.h:117-212).  Each
 * entry point below names the reference interface it replaces; the header-only C++ policy classes
 * in hemelb_b200/host/ forward to these calls.
 *
 * Conventions: every function returns 0 on success and non-zero on error, with the message
 * available from hlb_gpu_last_error() (thread-local).  All pointers are HOST pointers to arrays
 * laid out exactly as the reference holds them; they are borrowed for the duration of the call
 * only.  The handle owns all device memory.  There is no CPU fallback: without a CUDA device
 * hlb_gpu_create fails.  One handle = one MPI rank of the reference = one GPU.
 */
#ifndef HEMELB_B200_H
#define HEMELB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hlb_gpu_handle* hlb_gpu_t;

/* build_info strings of the reference (CMake/HemeLbOptions.cmake:54-68) as enums */
enum { HLB_KERNEL_LBGK = 0, HLB_KERNEL_MRT = 1, HLB_KERNEL_TRT = 2 };
enum { HLB_WALL_SIMPLEBOUNCEBACK = 0, HLB_WALL_BFL = 1, HLB_WALL_GZS = 2 };
enum { HLB_IOLET_NASHZEROTHORDERPRESSURE = 0, HLB_IOLET_LADD = 1 };
/* MacroscopicPropertyCache members (Code/lb/MacroscopicPropertyCache.h:50-90) */
enum {
  HLB_CACHE_DENSITY = 1, HLB_CACHE_VELOCITY = 2, HLB_CACHE_WALL_SHEAR_STRESS = 4,
  HLB_CACHE_VON_MISES_STRESS = 8, HLB_CACHE_SHEAR_RATE = 16, HLB_CACHE_STRESS_TENSOR = 32,
  HLB_CACHE_TRACTION = 64, HLB_CACHE_TANGENTIAL_TRACTION = 128,
  /* not a cache: gather the hlb_gpu_monitor values inside the collide kernels of this step */
  HLB_CACHE_MONITOR = 256
};

typedef struct {
  int lattice;            /* 15, 19 or 27 (HEMELB_LATTICE) */
  int kernel;             /* HLB_KERNEL_* (HEMELB_KERNEL) */
  int wall;               /* HLB_WALL_* (HEMELB_WALL_BOUNDARY) */
  int inlet, outlet;      /* HLB_IOLET_* (HEMELB_INLET_BOUNDARY / HEMELB_OUTLET_BOUNDARY) */
  double tau;             /* LbmParameters::GetTau(), Code/lb/LbmParameters.h:34-39 */
  int device;             /* CUDA ordinal */
  int rank, nranks;       /* Domain::GetLocalRank(), communicator size */
  int64_t n_sites;        /* Domain::GetLocalFluidSiteCount() */
  int64_t mid_count[6];   /* Domain::GetMidDomainCollisionCount(t), t = 0..5 */
  int64_t edge_count[6];  /* Domain::GetDomainEdgeCollisionCount(t) */
  int64_t total_shared_fs;/* Domain::totalSharedFs */
  int n_neighbours;       /* Domain::neighbouringProcs.size() */
  int n_inlets, n_outlets;/* BoundaryValues::GetLocalIoletCount() of the two objects */
  int reorder;            /* 1: renumber sites internally into long z-runs inside each site range
                             (needs hlb_gpu_set_site_coords for EVERY site); the API keeps speaking
                             reference site ids.  0: keep the reference order on the device. */
} hlb_gpu_config;

/* Iolet descriptor: 16 doubles {kind (0 cosine pressure / 1 parabolic velocity), normal[3],
 * position[3], radius, maxSpeed, densityMean, densityAmp, phase, period, warmUpLength,
 * minimumSimulationDensity, reserved}  (Code/lb/iolets/InOutLet{,Cosine,ParabolicVelocity}.h) */
#define HLB_IOLET_RECORD_DOUBLES 16

const char* hlb_gpu_last_error(void);
int hlb_gpu_device_count(int* count);

/* ---- construction: what SimBuilder / LBM::InitCollisions hand the streamers
 *      (Code/configuration/SimBuilder.h:115-269, Code/lb/lb.hpp:75-124) */
int hlb_gpu_create(const hlb_gpu_config* cfg, hlb_gpu_t* out);
int hlb_gpu_destroy(hlb_gpu_t h);
/* Domain::neighbourIndices (Code/geometry/Domain.h:526), reference form: int64, site-major,
 * values site*Q+dir | N*Q (rubbish) | N*Q+1+k (send slot).  May be called in site chunks. */
int hlb_gpu_set_neighbour_indices(hlb_gpu_t h, int64_t first_site, int64_t n, const int64_t* idx);
/* SiteData per site (Code/geometry/SiteDataBare.h:66-74): wall / iolet intersection masks
 * (bit d-1 <-> direction d) and iolet id */
int hlb_gpu_set_site_data(hlb_gpu_t h, int64_t first_site, int64_t n, const uint32_t* wall_mask,
                          const uint32_t* iolet_mask, const int32_t* iolet_id);
/* Domain::distanceToWall, n*(Q-1) doubles (Code/geometry/Domain.h:406-409,519) */
int hlb_gpu_set_wall_distances(hlb_gpu_t h, int64_t first_site, int64_t n, const double* dist);
/* Domain::wallNormalAtSite, n*3 doubles */
int hlb_gpu_set_wall_normals(hlb_gpu_t h, int64_t first_site, int64_t n, const double* normals);
/* Domain::globalSiteCoords, n*3 int64 */
int hlb_gpu_set_site_coords(hlb_gpu_t h, int64_t first_site, int64_t n, const int64_t* coords);
/* Domain::neighbouringProcs {Rank, SharedDistributionCount, FirstSharedDistribution}
 * (Code/geometry/NeighbouringProcessor.h:24-28) and streamingIndicesForReceivedDistributions */
int hlb_gpu_set_neighbours(hlb_gpu_t h, const int* rank, const int64_t* count, const int64_t* first);
int hlb_gpu_set_streaming_indices(hlb_gpu_t h, const int64_t* idx);
/* iolets of the inlet (which = 0) / outlet (which = 1) BoundaryValues object */
int hlb_gpu_set_iolets(hlb_gpu_t h, int which, int n, const double* records);
/* GZS site halo (geometry::neighbouring::NeighbouringDataManager, registered by
 * Code/lb/streamers/GuoZhengShi.h:93-99 and shared by NeighbouringDataManager::ShareNeeds):
 *  _remote: one entry per wall LINK whose GZS extrapolation reads the neighbour in `direction` (the
 *           inverse of the wall direction) of `local_site` when that neighbour lives on
 *           `owner_rank`; `owner_site` identifies the neighbour among that rank's sites (any id both
 *           sides can form: the reference's global non-contiguous site id, the owner's local id ...).
 *           Grouped by ascending owner rank.  Links with equal (owner_rank, owner_site) share ONE ghost
 *           row, as NeighbouringDataManager::RegisterNeededSite keeps a site once
 *           (NeighbouringDataManager.cc:27-39); ghost rows are numbered in order of first appearance.
 *  _serve:  the rows this rank SHIPS each step: (requester rank, local site), grouped by ascending
 *           requester rank and, within a rank, in the order of that requester's ghost rows (what
 *           NeighbouringDataManager::GetNeedsForProc(requester) lists). */
int hlb_gpu_set_gzs_remote(hlb_gpu_t h, int64_t n, const int64_t* local_site, const int32_t* direction,
                           const int32_t* owner_rank, const int64_t* owner_site);
int hlb_gpu_set_gzs_serve(hlb_gpu_t h, int64_t n, const int32_t* requester_rank, const int64_t* local_site);
int hlb_gpu_finalise(hlb_gpu_t h);

/* ---- multi-GPU: replaces net::Net's MPI point-to-point (Code/net/mixins/pointpoint/
 *      CoalescePointPoint.cc:23-132) with NCCL send/recv over NVLink */
int hlb_gpu_comm_unique_id(void* id128);
int hlb_gpu_comm_init(hlb_gpu_t h, const void* id128);

/* ---- FieldData (Code/geometry/FieldData.h:117-167, .cc:27-48) */
/* which = 0: f_old, 1: f_new; reference layout, N*Q + 1 + totalSharedFs doubles */
int hlb_gpu_set_f(hlb_gpu_t h, int which, const double* f);
int hlb_gpu_get_f(hlb_gpu_t h, int which, double* f);
/* the halo region alone (totalSharedFs doubles after the rubbish slot): lets a host keep the
 * reference's own net::Net / MPI for the exchange (host-staged) instead of NCCL.  get: the send
 * slices of f_new (which = 1); set: the receive slices of f_old (which = 0). */
int hlb_gpu_get_halo(hlb_gpu_t h, int which, double* out);
int hlb_gpu_set_halo(hlb_gpu_t h, int which, const double* in);
/* EquilibriumInitialCondition::SetFs (Code/lb/InitialCondition.hpp:40-52) */
int hlb_gpu_set_equilibrium(hlb_gpu_t h, double density, const double* momentum3);
/* phase 0 of a step: NeighbouringDataManager::TransferFieldDependentInformation
 * (NeighbouringDataManager.cc:101-142) -- pack and ship the whole f_old rows of the _serve list,
 * receive the _remote ones (NCCL; or stage them through the host with the two calls below) */
int hlb_gpu_exchange_site_halo(hlb_gpu_t h);
int hlb_gpu_get_gzs_send(hlb_gpu_t h, double* out);     /* n_serve * Q doubles, site-major */
int hlb_gpu_set_gzs_ghost(hlb_gpu_t h, const double* in); /* n_remote * Q doubles, site-major */
int hlb_gpu_request_comms(hlb_gpu_t h);   /* FieldData::SendAndReceive + Net::Receive/Send */
int hlb_gpu_copy_received(hlb_gpu_t h);   /* Net::Wait + FieldData::CopyReceived */
int hlb_gpu_swap(hlb_gpu_t h);            /* FieldData::SwapOldAndNew */

/* ---- per-step scalars: SimulationState::GetTimeStep() (1-indexed), BoundaryValues::
 *      GetBoundaryDensity(i) for each inlet / outlet, and the refresh flags of the property cache
 *      (Code/SimulationMaster.impl.h:223-241) */
int hlb_gpu_set_step_scalars(hlb_gpu_t h, uint64_t time_step, const double* inlet_density,
                             const double* outlet_density, uint32_t cache_mask);

/* ---- lb::streamer concept: slot 0..5 = mid-fluid, wall, inlet, outlet, inlet-wall,
 *      outlet-wall streamer of LBM (Code/lb/lb.h:102-107); (first, count) as LBM passes them */
int hlb_gpu_stream_and_collide(hlb_gpu_t h, int slot, int64_t first, int64_t count);
int hlb_gpu_post_step(hlb_gpu_t h, int slot, int64_t first, int64_t count);
/* mark the end of LBM::PreSend (all domain-edge ranges issued): lets the halo send start while
 * the mid-domain ranges run */
int hlb_gpu_edge_done(hlb_gpu_t h);

/* ---- MacroscopicPropertyCache read-back (site-major doubles as the reference's caches) */
int hlb_gpu_get_cache(hlb_gpu_t h, uint32_t which, double* out);

/* ---- whole steps: the phase sequence of LBM (RequestComms, PreSend, PreReceive, PostReceive,
 *      EndIteration; Code/lb/lb.hpp:162-314) + SwapOldAndNew + SimulationState::Increment, with
 *      cosine iolet densities evaluated on the host as InOutLetCosine::GetDensity does */
int hlb_gpu_step(hlb_gpu_t h, int nsteps);
int hlb_gpu_get_time_step(hlb_gpu_t h, uint64_t* t);
/* scheduling knob.  Product schedule (enabled = 1, as created): requests for a whole range with that
 * range's own streamer (what LBM::PreSend / PreReceive make, lb.hpp:176-251) are held and served by ONE
 * launch of the site kernel per part (domain-edge, mid-domain) in which every site runs the streamer of
 * its own collision type; the twelve PostStep requests become one launch over the BFL link list
 * (INTEGRATION.md "Asynchrony").  The ranges read f_old and write disjoint slots of
 * f_new, so the result of a step is bit-identical in every schedule; whatever reads or orders state
 * (CopyReceived, swap, read-backs, monitors) first sends off what was held.  enabled = 0: every request
 * its own launch, in the caller's order (A/B comparisons).  Sub-range requests, and a streamer asked to
 * run on another type's range, always run at once.  Environment at create: HLB_SCHEDULE=0 (as enabled = 0),
 * HLB_PREFETCH=<sites>, HLB_NBR_RUNS=0, HLB_TMA=1, HLB_GZS_OVERLAP=k (measurement switches). */
int hlb_gpu_set_overlap(hlb_gpu_t h, int enabled);
int hlb_gpu_sync(hlb_gpu_t h);
/* CUDA-event timing of nsteps whole steps on the engine's own streams (ms) */
int hlb_gpu_time_steps(hlb_gpu_t h, int nsteps, float* ms);
/* same, also returning the summed CUDA-event duration of the site-kernel launches over the mid-domain
 * part and the number of sites they updated -- the roofline kernel, timed live inside the step */
int hlb_gpu_time_steps_detail(hlb_gpu_t h, int nsteps, float* total_ms, float* bulk_ms, int64_t* bulk_sites);
/* device-side monitors (StabilityTester / IncompressibilityChecker inputs): {min f, min density,
 * max density, max |u|}; D2H of 4 doubles.  With HLB_CACHE_MONITOR in the step's cache mask the
 * values were gathered by the collide kernels from the distributions ENTERING the last step(s)
 * since the previous call (no extra pass); otherwise one pass over the current f_old. */
int hlb_gpu_monitor(hlb_gpu_t h, double* out4);
/* the same read-back in two halves: _begin enqueues the reduction and the 32-byte copy behind the step
 * just issued and returns at once; _end waits for them and hands the values out.  Between the two the
 * caller may issue the next time step, so that the device never idles while the host looks at the
 * monitors (the values are then one step old; the reference's StabilityTester / IncompressibilityChecker
 * cycles span several time steps, Code/net/PhasedBroadcast.h).  One read-back outstanding at a time. */
int hlb_gpu_monitor_begin(hlb_gpu_t h);
int hlb_gpu_monitor_end(hlb_gpu_t h, double* out4);
/* the same four values over ALL ranks: one ncclAllReduce on the handle's communicator
 * (hlb_gpu_comm_init) in place of the PhasedBroadcast trees of lb::StabilityTester
 * (Code/lb/StabilityTester.h:51-150) and lb::IncompressibilityChecker
 * (Code/lb/IncompressibilityChecker.hpp); collective -- every rank calls it at the same point of the
 * step sequence.  One rank: same as hlb_gpu_monitor. */
int hlb_gpu_monitor_global(hlb_gpu_t h, double* out4);
/* number of kernels launched by this handle so far */
int hlb_gpu_launch_count(hlb_gpu_t h, int64_t* n);
/* How much of geometry::Domain's neighbourIndices (Code/geometry/Domain.cc:425-505) the whole-part
 * launches read in run form: of the `words` groups of 32 consecutive device sites, `words_in_runs`
 * have every streaming target given by at most two runs per direction (8 B per direction and group
 * instead of 128 B).  0 of 0 with HLB_NBR_RUNS=0.  After hlb_gpu_finalise. */
int hlb_gpu_target_runs(hlb_gpu_t h, int64_t* words_in_runs, int64_t* words);
/* lb::StabilityTester<LATTICE>::PostSendToParent's loop over the local sites
 * (Code/lb/StabilityTester.h:97-141) as one device reduction, to be called where that loop runs:
 * after the step's streaming, before SwapOldAndNew.  out2[0] = the number of populations of f_new
 * that fail "value > 0.0" (zero, negative or NaN: Unstable when non-zero); out2[1] = the largest
 * |u_new - u_old| over the local sites when with_convergence is non-zero (the absolute error of
 * ComputeRelativeDifference, :156-180: relative difference = out2[1] / convergenceReferenceValue), else 0.
 * 16 bytes leave the device instead of the two whole distribution arrays. */
int hlb_gpu_stability(hlb_gpu_t h, int with_convergence, double* out2);

/* the device-resident tables, converted back to reference form (for parity tests) */
int hlb_gpu_get_neighbour_indices(hlb_gpu_t h, int64_t* idx);

/* ==== device-side geometry::Domain construction ==================================================
 * Replaces, for one rank, the host loops of geometry::Domain (Code/geometry/Domain.cc:69-357 site
 * order and collision-type buckets, :425-505 neighbourIndices, :247-285 neighbouringProcs,
 * :507-580 halo slots and streamingIndicesForReceivedDistributions) fed by GeometryReader
 * (Code/geometry/GeometryReader.cc:556-652).  Every table it produces is bit-identical to the
 * reference's; it can be read back in reference form (parity tests) or handed to an engine handle
 * device-to-device (hlb_gpu_create_from_domain), so the N*Q int64 table never visits the host.
 * Two site sources: an explicit .gmy-level site list, or an analytic union of capsules clipped by
 * flat iolet caps voxelised on the device (synthetic geometries; the link model of
 * doc/dev/file-formats/geometry.md: per-link cut type / iolet id / float32 distance, wall normal). */
typedef struct hlb_dom_handle* hlb_dom_t;

typedef struct {
  int lattice;            /* 15, 19 or 27 */
  int block_size;         /* sites per block side (.gmy header), <= 8 */
  int64_t block_dims[3];  /* blocks per axis (.gmy header) */
  int rank, nranks;       /* the rank whose Domain is built */
  int device;             /* CUDA ordinal */
} hlb_dom_config;

int hlb_dom_create(const hlb_dom_config* cfg, hlb_dom_t* out);
int hlb_dom_destroy(hlb_dom_t d);
/* source A: fluid sites as GeometryReader delivers them.  coords: n_sites x 3 global voxel
 * coordinates (any order; only this rank's sites and their lattice neighbours are needed);
 * rank_of_site: n_sites (NULL = all on rank 0) -- the decomposition's answer, ParMETIS or basic;
 * cut-link records for the sites that have any: record_site[n_records] indexes coords, link arrays
 * are n_records x 26 in the file's neighbour order (Code/io/formats/geometry.h:120-156). */
int hlb_dom_set_sites(hlb_dom_t d, int64_t n_sites, const int32_t* coords, const int32_t* rank_of_site,
                      int64_t n_records, const int64_t* record_site, const uint8_t* link_type,
                      const int32_t* link_iolet, const float* link_dist, const uint8_t* normal_available,
                      const float* normal);
/* source B: union of capsules {a[3], b[3], radius} (7 doubles each) clipped by iolet caps
 * {kind (2 inlet / 3 outlet), index, position[3], normal[3] (into the fluid), radius} (9 doubles
 * each), in voxel units of the block lattice. */
int hlb_dom_set_shape(hlb_dom_t d, int n_capsules, const double* capsules, int n_iolets, const double* iolets);
/* source B, optional: wall roughness.  The surface of capsule k is displaced by amplitude[k] x noise(p),
 * noise = trilinear interpolation of grid^3 seeded values in [-1, 1] (x-major) stretched over
 * extent[3] voxels (the lattice); cut distances by bisection on the displaced implicit function, wall
 * normals by central differences (hemelb_b200/geometry.py: sac, voxelise -- configs[4], the
 * aneurysm-like sac = one rough sphere + a neck cylinder).  After hlb_dom_set_shape. */
int hlb_dom_set_roughness(hlb_dom_t d, const double* amplitude, int grid, const double* noise, const double* extent);
/* site -> rank rule for source B: slabs along an axis (rank r owns first_coord[r] <= x < first_coord
 * [r+1]; nranks + 1 ascending values; cuts through blocks like a ParMETIS site partition), or
 * whole blocks (rank_of_block[prod(block_dims)], .gmy block order; what BasicDecomposition,
 * Code/geometry/decomposition/BasicDecomposition.cc:21-96, produces) */
int hlb_dom_set_partition_slabs(hlb_dom_t d, int axis, const int64_t* first_coord);
int hlb_dom_set_partition_blocks(hlb_dom_t d, const int32_t* rank_of_block);
/* source B: fluid sites per block over a box of blocks [lo, hi) (the input of BasicDecomposition);
 * counts is x-major / z-fastest over the box */
int hlb_dom_count_block_sites(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts);
/* same, plus how many of each block's sites are boundary-typed (some lattice link is cut by a wall
 * or an iolet; Domain.cc:186-207) -- the vertex weights of a weighted decomposition
 * (DecompositionWeights.h.in:25-62, OptimisedDecomposition::PopulateVertexWeightData) */
int hlb_dom_count_block_sites_typed(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts,
                                    int32_t* boundary_counts);
int hlb_dom_build(hlb_dom_t d);
int hlb_dom_build_seconds(hlb_dom_t d, double* seconds);  /* device time of the last build */
/* ---- the tables, reference form */
int hlb_dom_get_counts(hlb_dom_t d, int64_t* n_sites, int64_t* mid6, int64_t* edge6, int64_t* total_shared_fs,
                       int* n_neighbours);
int hlb_dom_get_neighbours(hlb_dom_t d, int* rank, int64_t* count, int64_t* first);
int hlb_dom_get_streaming_indices(hlb_dom_t d, int64_t* idx);
int hlb_dom_get_neighbour_indices(hlb_dom_t d, int64_t first_site, int64_t n, int64_t* idx);
int hlb_dom_get_site_coords(hlb_dom_t d, int64_t first_site, int64_t n, int64_t* coords);  /* n x 3 */
int hlb_dom_get_input_index(hlb_dom_t d, int64_t first_site, int64_t n, int64_t* idx);     /* source A */
/* the boundary-typed sites (local ids [mid[0], sum(mid)) then [sum(mid)+edge[0], N)), site-major:
 * SiteData masks + iolet id, distanceToWall (Q-1 per site, -1 where uncut), wallNormalAtSite
 * (3 per site, +inf where the file has none) */
int hlb_dom_get_boundary_tables(hlb_dom_t d, uint32_t* wall_mask, uint32_t* iolet_mask, int32_t* iolet_id,
                                double* dist, double* normal);
/* GuoZhengShi across ranks: what the constructor loop of GuoZhengShiLink registers with the
 * NeighbouringDataManager (Code/lb/streamers/GuoZhengShi.h:36-104) -- for every wall link of a local site
 * whose opposite direction is neither wall nor iolet and leads to a site on another rank: the local site,
 * that opposite direction, the owning rank and the neighbour's global coordinates -- ordered by owner
 * rank, site, direction (the rows of hlb_gpu_set_gzs_remote).  *n = how many there are; the arrays are
 * filled when capacity >= *n.  hlb_dom_lookup_sites: local site ids of global coordinates (-1: not a
 * local fluid site) -- how the owner turns the coordinates it is asked for into its serve list. */
int hlb_dom_gzs_needs(hlb_dom_t d, int64_t capacity, int64_t* n, int64_t* local_site, int32_t* direction,
                      int32_t* owner_rank, int64_t* coords);
int hlb_dom_lookup_sites(hlb_dom_t d, int64_t n, const int64_t* coords, int64_t* local_site);
/* source B: the voxelised geometry in .gmy terms (sites in traversal order; one record per site
 * with a non-fluid 26-neighbour), e.g. to write a .gmy the reference can read */
int hlb_dom_get_geometry_sizes(hlb_dom_t d, int64_t* n_sites, int64_t* n_records);
int hlb_dom_get_geometry(hlb_dom_t d, int32_t* coords, int64_t* record_site, uint8_t* link_type, int32_t* link_iolet,
                         float* link_dist, uint8_t* normal_available, float* normal);
/* hlb_gpu_create + every hlb_gpu_set_* table call, device-to-device.  `policy` supplies kernel,
 * wall, inlet, outlet, tau, n_inlets, n_outlets, reorder; sizes, ranks and device come from the
 * built domain.  Still to do on the handle: hlb_gpu_set_iolets, hlb_gpu_finalise. */
int hlb_gpu_create_from_domain(hlb_dom_t d, const hlb_gpu_config* policy, hlb_gpu_t* out);

/* ==== property extraction (.xtr) and checkpoints ==================================================
 * Replaces, for one rank, the per-site loops of extraction::LocalPropertyOutput::Write
 * (Code/extraction/LocalPropertyOutput.cc:262-367) over extraction::LbDataSourceIterator
 * (Code/extraction/LbDataSourceIterator.cc:36-87: property cache -> physical units through
 * util::UnitConverter) and the geometry selectors (Code/extraction/{Whole,Plane,StraightLine,...}GeometrySelector.cc), and the
 * record decoding of extraction::LocalDistributionInput::LoadDistribution
 * (Code/extraction/LocalDistributionInput.cc:107-165).  The bytes are the reference's file format
 * (doc/dev/file-formats/extraction.md, version 5, XDR big-endian) exactly.  File offsets across
 * ranks (the Scan/AllReduce of LocalPropertyOutput.cc:96-110), the .off file and the actual
 * MPI-IO / POSIX writes stay with the host (hemelb_b200/extraction.py; in a HemeLB build the
 * unchanged LocalPropertyOutput constructor) -- the handle produces the header and this rank's
 * record bytes. */
typedef struct hlb_xtr_handle* hlb_xtr_t;
/* extraction::source::Type (Code/extraction/OutputField.h:18-40) */
enum {
  HLB_XTR_PRESSURE = 0, HLB_XTR_VELOCITY = 1, HLB_XTR_SHEARSTRESS = 2, HLB_XTR_VONMISESSTRESS = 3,
  HLB_XTR_SHEARRATE = 4, HLB_XTR_STRESSTENSOR = 5, HLB_XTR_TRACTION = 6,
  HLB_XTR_TANGENTIALPROJECTIONTRACTION = 7, HLB_XTR_DISTRIBUTIONS = 8, HLB_XTR_MPIRANK = 9
};
/* io::formats::extraction::TypeCode (Code/io/formats/extraction.h:25-32) */
enum { HLB_XTR_FLOAT = 0, HLB_XTR_DOUBLE = 1, HLB_XTR_INT32 = 2, HLB_XTR_UINT32 = 3, HLB_XTR_INT64 = 4, HLB_XTR_UINT64 = 5 };
/* GeometrySelector subclasses; selector_params: plane {point[3], normal[3], radius (<= 0: infinite)},
 * line {endpoint1[3], endpoint2[3]}, surface point {point[3]} -- physical units, float as the reference */
enum {
  HLB_XTR_WHOLE = 0, HLB_XTR_SURFACE = 1, HLB_XTR_PLANE = 2, HLB_XTR_LINE = 3, HLB_XTR_SURFACEPOINT = 4,
  /* plane whose normal is used as given: what PlaneGeometrySelector::GetNormal() returns (the
   * constructor normalised it already, PlaneGeometrySelector.cc:12-29); HLB_XTR_PLANE takes the
   * configured normal and normalises it exactly as that constructor does */
  HLB_XTR_PLANE_NORMALISED = 5
};

typedef struct {          /* extraction::OutputField (Code/extraction/OutputField.h:105-112) */
  const char* name;
  int source;             /* HLB_XTR_PRESSURE ... */
  int typecode;           /* HLB_XTR_FLOAT ... */
  uint32_t n_offsets;     /* 0, 1 or the field length; only Pressure subtracts offsets[0] from the data
                             (LocalPropertyOutput.cc:308-310), the rest is header information */
  const double* offsets;
} hlb_xtr_field;

typedef struct {          /* extraction::PropertyOutputFile + the util::UnitConverter constructor arguments */
  int selector;           /* HLB_XTR_WHOLE ... */
  float selector_params[7];
  int n_fields;           /* <= 16 */
  const hlb_xtr_field* fields;
  double time_step, voxel_size, origin[3], fluid_density, reference_pressure;
} hlb_xtr_spec;

/* evaluates the selector for every local site on the device (CountWrittenSitesOnRank,
 * LocalPropertyOutput.cc:133-145) and keeps the included-site list.  site_coords:
 * Domain::globalSiteCoords, n_sites x 3 int64 (borrowed for the call).  `h` must be finalised. */
int hlb_xtr_create(hlb_gpu_t h, const hlb_xtr_spec* spec, const int64_t* site_coords, hlb_xtr_t* out);
/* same with the coordinates of a device-built Domain (no host copy) */
int hlb_xtr_create_from_domain(hlb_gpu_t h, hlb_dom_t d, const hlb_xtr_spec* spec, hlb_xtr_t* out);
int hlb_xtr_destroy(hlb_xtr_t x);
/* local_site_count; bytes per site record (CalcSiteWriteLen); MainHeaderLength + field header length */
int hlb_xtr_sizes(hlb_xtr_t x, uint64_t* local_site_count, uint64_t* site_length, uint64_t* header_length);
/* PropertyActor::SetRequiredProperties (Code/extraction/PropertyActor.cc:22-75): HLB_CACHE_* the fields need */
int hlb_xtr_required_caches(hlb_xtr_t x, uint32_t* cache_mask);
/* LocalPropertyOutput::PrepareHeader (LocalPropertyOutput.cc:178-213): main header + field headers */
int hlb_xtr_header(hlb_xtr_t x, uint64_t global_site_count, void* buf, uint64_t capacity);
/* the record bytes of included sites [first_site, first_site + n_sites) of the current state (property
 * caches of the last step, f_old), encoded on the device and copied to host_buf (pageable or pinned).
 * Slices let a large checkpoint stream through a bounded buffer. */
int hlb_xtr_encode(hlb_xtr_t x, uint64_t first_site, uint64_t n_sites, void* host_buf, uint64_t capacity);
/* a pinned host buffer owned by the handle (grown on demand) for hlb_xtr_encode to land in */
int hlb_xtr_pinned_buffer(hlb_xtr_t x, uint64_t bytes, void** ptr);
/* device time of the last hlb_xtr_encode's kernel (CUDA events on the engine's stream) */
int hlb_xtr_last_encode_ms(hlb_xtr_t x, float* ms);
/* checkpoint: this rank's slice of one time step of a distributions-only double extraction file
 * (without the IO rank's leading 8-byte time stamp).  Checks every record's grid position against the
 * local site at that index and sets f_old = f_new = the stored values. */
int hlb_gpu_load_distributions(hlb_gpu_t h, const void* records, uint64_t n_bytes, const int64_t* site_coords);
int hlb_gpu_load_distributions_from_domain(hlb_gpu_t h, hlb_dom_t d, const void* records, uint64_t n_bytes);

/* ==== METIS-free k-way partition of the site graph (host code) ====================
 * Stands where the reference calls ParMETIS_V3_PartKway (Code/geometry/decomposition/
 * OptimisedDecomposition.cc:138-154), on one process: the CSR graph of PopulateAdjacencyData
 * (:311-379; xadj[n + 1], adjncy[xadj[n]], symmetric), vertex weights by collision type
 * (DecompositionWeights.h.in:25-62, integer-valued), nparts, ubvec (:133).  `part` holds the
 * partition the vertices arrive with (BasicDecomposition's, or hlb_part_bisect's) and is refined in
 * place: parts above max(ubvec x mean, mean + heaviest vertex) diffuse boundary vertices along flows
 * solved on the part graph; then boundary vertices move, in bulk sweeps, to the part they have most
 * links into while the number of cut links falls.  Deterministic.  edgecut (may be NULL): cut links
 * of the result, as ParMETIS reports them. */
int hlb_part_refine_kway(int64_t n, const int64_t* xadj, const int64_t* adjncy, const double* vwgt, int nparts,
                         double ubvec, int passes, int32_t* part, int64_t* edgecut);
/* recursive geometric bisection of weighted points (coords: n x 3 voxel or block coordinates) into
 * nparts, floor(k/2) : k - floor(k/2) as BasicDecomposition.cc:21-96 divides its ranks: at the
 * weighted median across the longest coordinate extent (inertial = 0) or along the principal axis
 * of the weighted covariance (inertial = 1) -- cuts vessels across rather than along the Morton
 * curve. */
int hlb_part_bisect(int64_t n, const int64_t* coords, const double* weights, int nparts, int inertial, int32_t* part);

#ifdef __cplusplus
}
#endif
#endif /* HEMELB_B200_H */
