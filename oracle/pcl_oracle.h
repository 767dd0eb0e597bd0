/*
 * pcl_oracle.h — C API of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the upstream PCL 1.8.x / FLANN
 * semantics that the reference programs (/root/reference/SHOT.cpp, SHOT_demo.cpp, FPFH_demo.cpp,
 * 6Dpose.cpp, CAD_desc.cpp ...) reach through their PCL calls.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product (libb200reg.so) never
 * links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference repo holds no tests, golden vectors or data files, and PCL itself is
 * not vendored nor installed here.  The algorithm is restated from the published
 * PCL 1.8 / FLANN 1.8 sources as recorded in SURVEY.md Appendix A; every function below cites the
 * reference call site (file:line under /root/reference) whose behaviour it stands in for and the
 * upstream PCL file it follows.
 *
 * Conventions: points are `float` rows with a caller-given stride (in floats, >= 3; x,y,z first).
 * Normals are rows of 4 floats (nx, ny, nz, curvature).  All squared distances are float32 sums
 * dx*dx + dy*dy + dz*dz evaluated left to right with no fused multiply-add (FLANN L2_Simple).
 */
#ifndef PCL_ORACLE_H_
#define PCL_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int index_query;  /* model descriptor / keypoint index  (pcl::Correspondence::index_query) */
  int index_match;  /* scene descriptor / keypoint index  (pcl::Correspondence::index_match) */
  float distance;   /* squared L2 descriptor distance */
} orc_corr;

/* Sensitivity variants (bit mask, 0 = the documented restatement): 1 Eigen 4-lane dot order, 2 float32 Umeyama
 * moments, 4 degenerate FPFH pairs skipped, 8 / 16 FPFH f1 bin coordinate +- 1e-6, 32 / 64 eigen33 theta +- 2 ulp,
 * 128 / 256 SHOT cosine-bin coordinate +- 1e-6, 512 / 1024 BOARD direction angle +- 2 ulp, 2048 hypothesis verification:
 * the explaining model point of a scene point is the closest one (hv_oracle.cpp).  See oracle/sensitivity.py and
 * oracle/sensitivity_hv.py. */
void orc_set_variant(unsigned mask);
unsigned orc_get_variant(void);
int orc_num_threads(void);
void orc_set_num_threads(int n);

/* KdTreeFLANN::radiusSearch (SHOT_VAR.cpp:356; implicit in every Feature::compute with
 * setRadiusSearch).  Neighbours with d2 < (float)(radius*radius), sorted by (d2, index).
 * CSR output: offsets[nq+1]; idx/d2 hold up to `cap` entries.  Returns the total number of
 * neighbours (call once with cap = 0 to size the buffers; offsets is always filled). */
int64_t orc_radius_search(const float *surf, int n, int sstride, const float *q, int nq, int qstride,
                          double radius, int64_t *offsets, int *idx, float *d2, int64_t cap);

/* KdTreeFLANN::nearestKSearch (SHOT.cpp:163, Edge_detection.cpp:120; implicit with setKSearch).
 * k is clamped to the number of finite surface points; rows are sorted by (d2, index); unused
 * slots are filled with -1 / +inf.  Returns the clamped k. */
int orc_knn_search(const float *surf, int n, int sstride, const float *q, int nq, int qstride, int k,
                   int *idx, float *d2);

/* NormalEstimationOMP::compute with setKSearch(k) (SHOT.cpp:302-308, 6Dpose.cpp:275-278,
 * SHOT_demo.cpp:405-411) or setRadiusSearch(r) (FPFH_demo.cpp:416-420).  Exactly one of k / radius
 * must be non-zero.  Queries = input cloud, surf = search surface (pass the same array for both, as
 * every reference call site does).  out: nq rows of 4 floats. */
int orc_normals(const float *surf, int n, int sstride, const float *q, int nq, int qstride, int k,
                double radius, const float *viewpoint3, float *out);

/* SHOTLocalReferenceFrameEstimationOMP (built implicitly by SHOTEstimationOMP::initCompute,
 * SHOT.cpp:360-366).  out: K rows of 9 floats (x axis, y axis, z axis); NaN rows when < 5 valid
 * neighbours. */
int orc_shot_lrf(const float *surf, int n, int sstride, const float *kp, int K, int kstride,
                 double radius, float *out);

/* SHOTEstimationOMP<.., SHOT352>::compute (SHOT.cpp:360-371, SHOT_demo.cpp:419-424, 497-502,
 * CAD_desc.cpp:341-352).  normals: n rows of 4 floats for the search surface.
 * desc: K x 352, rf: K x 9. */
int orc_shot352(const float *surf, const float *normals, int n, int sstride, const float *kp, int K,
                int kstride, double radius, float *desc, float *rf);

/* FPFHEstimation / FPFHEstimationOMP ::compute (FPFH_demo.cpp:422-428, 505-510;
 * FPFH_scenes_clustered.cpp:287-293).  q == NULL means input == surface (every reference call
 * site).  out: nq (or n) rows of 33 floats. */
int orc_fpfh33(const float *surf, const float *normals, int n, int sstride, const float *q, int nq,
               int qstride, double radius, float *out);

/* KdTreeFLANN<Descriptor>::setInputCloud + the user matching loop.
 * mode 1: k = 1, accept iff d2 < thr            (SHOT.cpp:405-423, SHOT_scenes.cpp:359-365)
 * mode 2: k = 2, accept iff d0/d1 <= 1 (float)   (SHOT_demo.cpp:508-530, FPFH_demo.cpp:516-538)
 * Returns the number of correspondences written to out (capacity Ks). */
int orc_match(const float *model, int Km, const float *scene, int Ks, int D, int mode, float thr,
              orc_corr *out);
/* OpenMP-parallel variant of the same loop (labelled baseline only; the reference loop is serial). */
int orc_match_omp(const float *model, int Km, const float *scene, int Ks, int D, int mode, float thr,
                  orc_corr *out);

/* GeometricConsistencyGrouping::recognize (SHOT.cpp:473-482, 6Dpose.cpp:529-538).
 * transforms: up to max_inst x 16 floats (row-major 4x4, model -> scene).
 * inst_offsets: max_inst + 1; inst_corrs: capacity corr_cap.  Returns the number of instances
 * (all are counted even if the buffers are too small; nothing is written past the capacities). */
int orc_gc_recognize(const float *model_kp, int Km, int mstride, const float *scene_kp, int Ks,
                     int sstride, const orc_corr *corrs, int C, double gc_size, int gc_threshold,
                     float *transforms, int max_inst, int *inst_offsets, orc_corr *inst_corrs,
                     int corr_cap);

/* Building blocks exposed for unit tests. */
void orc_eigen33_smallest(const float cov9[9], float *eigenvalue, float evec3[3]); /* pcl::eigen33 */
void orc_eigh3_f64(const double a9[9], double evals3[3], double evecs9[9]);          /* ascending; columns */
void orc_umeyama3(const double *src, const double *dst, int n, double T16[16]);      /* rigid, no scale */
uint32_t orc_mt19937_nth(uint32_t seed, int nth);                                    /* known-answer hook */

/* pcl::Hough3DGrouping::recognize with given reference frames, no interpolation, distance weight on (the
 * reference's default grouping, SHOT.cpp:433-470).  rf: K x 9 floats.  Outputs as orc_gc_recognize. */
int orc_hough3d_recognize(const float *model_kp, const float *model_rf, int Km, int mstride, const float *scene_kp,
                          const float *scene_rf, int Ks, int sstride, const orc_corr *corrs, int C, double bin_size,
                          double threshold, float *transforms, int max_inst, int *inst_offsets, orc_corr *inst_corrs,
                          int corr_cap);

/* pcl::BOARDLocalReferenceFrameEstimation (SHOT.cpp:441-453, 6Dpose.cpp:497-509, FPFH_demo.cpp:556-568).  Defaults of
 * the PCL constructor: tangent_radius 0, find_holes false (the reference sets true), margin_thresh 0.85,
 * check_margin_array_size 24, hole_size_prob_thresh 0.2, steep_thresh 0.1.  out: K x 9 (x, y, z axes); returns the
 * number of rand() values drawn (glibc generator seeded rand_seed, rand_skip values skipped first). */
typedef struct {
  int find_holes;
  float tangent_radius;
  float margin_thresh;
  int check_margin_array_size;
  float hole_size_prob_thresh;
  float steep_thresh;
} orc_board_params;
int orc_board_lrf(const float *surf, const float *normals, int n, int sstride, const float *kp, int K, int kstride,
                  double radius, const orc_board_params *bp, unsigned rand_seed, int rand_skip, float *out);
int orc_glibc_rand_nth(unsigned seed, int nth); /* nth value of rand() after srand(seed), nth >= 1 */

/* pcl::removeNaNFromPointCloud / pcl::transformPointCloud (SHOT.cpp:298-299; the model placed by a pose before ICP). */
int orc_remove_nan(const float *xyz, int n, int stride, float *out_xyz, int *out_index);
void orc_transform_points(const float *xyz, int n, int stride, const float *T16, float *out_xyz);

/* pcl::IterativeClosestPoint::align + getFitnessScore (SHOT.cpp:177-192, SHOT_demo.cpp:604-663).  PCL defaults:
 * max_corr_dist <= 0 = unlimited, transformation_epsilon 0, euclidean_fitness_epsilon -DBL_MAX.  final_T: row-major
 * 4x4; aligned (nullable): ns x 3. */
int orc_icp_align(const float *source, int ns, int sstride, const float *target, int nt, int tstride, int max_iterations,
                  double max_corr_dist, double transformation_epsilon, double euclidean_fitness_epsilon,
                  const float *guess, float *final_T, float *aligned, double *fitness, int *converged, int *iterations);

/* pcl::UniformSampling::filter (SHOT.cpp:314-323) / pcl::VoxelGrid::filter (SHOT_demo.cpp:413-417); output
 * in ascending leaf index; out_xyz has room for n x 3 floats.  Return the number of keypoints, -1 when the
 * lattice is too fine (PCL: "leaf size is too small").  See the implementation for the PCL semantics. */
int orc_uniform_sampling(const float *xyz, int n, int stride, double leaf, float *out_xyz, int *out_index);
int orc_voxel_grid(const float *xyz, int n, int stride, float lx, float ly, float lz, float *out_xyz);

/* pcl::GlobalHypothesesVerification as SHOT_hypothesis.cpp:631-653 drives it (setSceneCloud, addModels(.., true), the
 * set* calls, verify, getMask); restated in hv_oracle.cpp — see its header for what could not be checked.
 * occlusion_threshold is the value in force when addModels ran (the reference calls setOcclusionThreshold AFTER
 * addModels, so its models are filtered with the constructor default 0.005).  detect_clutter must be 0. */
typedef struct {
  float resolution;              /* HypothesisVerification::resolution_ 0.005 (voxel size of scene and hypotheses) */
  float inlier_threshold;        /* setInlierThreshold (hv_inlier_th_ 0.005, SHOT_hypothesis.cpp:59) */
  float occlusion_threshold;     /* occlusion_thres_ at addModels time */
  float regularizer;             /* setRegularizer (hv_regularizer_) */
  float radius_normals;          /* setRadiusNormals (hv_rad_normals_) */
  float res_occupancy_grid;      /* 0.01 */
  float w_occupied_multiple_cm;  /* 4 */
  float initial_temp;            /* 1000 */
  int max_iterations;            /* 5000 (noimprove_termination_criteria) */
  int occlusion_reasoning;       /* addModels' second argument */
  int zbuffer_scene_resolution;  /* 100 */
  int zbuffer_self_resolution;   /* 75 */
  float self_occlusion_threshold; /* 0.005 */
  int detect_clutter;            /* setDetectClutter; only 0 is restated */
  float radius_clutter;          /* unused without the clutter cue */
  float clutter_regularizer;     /* unused without the clutter cue */
  unsigned rand_seed;            /* srand() state std::random_shuffle draws from (1 = never seeded) */
  unsigned mt_seed;              /* mt19937 default seed 5489 */
  int sa_uniform_mode;           /* 0: x / 2^32 (boost::uniform_real), 1: the raw 32-bit value (tr1 on a bare engine) */
} orc_hv_params;
typedef struct {
  int valid;        /* addModel returned true */
  int n_visible;    /* points left by the occlusion filters */
  int n_points;     /* after VoxelGrid and the NaN-normal compaction */
  int n_outliers;   /* bad_information_ */
  int n_explained;  /* explained_.size() */
  int n_occupancy;  /* complete_cloud_occupancy_indices_.size() */
  float outliers_weight;
  float explained_sum; /* sequential float32 sum of explained_distances_ */
} orc_hv_info;
void orc_hv_params_default(orc_hv_params *p);
/* models: H clouds concatenated, model_offsets[H+1] in points.  mask: H bytes.  info: H entries (nullable).
 * Returns 0, -1 detect_clutter set, -2 voxel lattice too fine, -3 occupancy grid too large. */
int orc_hv_verify(const float *scene, int n, int sstride, const float *models, const int *model_offsets, int H, int mstride,
                  const orc_hv_params *P, unsigned char *mask, orc_hv_info *info, double *best_cost, int *accepted_moves,
                  int *n_scene_points, int *n_cells);
/* SAOptimize alone on given cue lists (CSR over the H hypotheses). */
int orc_hv_optimize(int H, int ns, const int *expl_off, const int *expl_idx, const float *expl_w, const int *occ_off,
                    const int *occ_idx, int n_cells, const float *outliers_weight, const int *bad_information,
                    const orc_hv_params *P, unsigned char *mask, double *best_cost, int *accepted_moves);
/* cue lists of the last orc_hv_verify on this thread: 0 expl_off, 1 expl_idx, 2 expl_w (float), 3 occ_off, 4 occ_idx */
int orc_hv_last_size(int which);
int orc_hv_last_copy(int which, void *dst);

#ifdef __cplusplus
}
#endif
#endif
