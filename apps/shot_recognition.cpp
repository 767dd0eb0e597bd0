// shot_recognition.cpp — the recognition callback of the reference, driven through the PCL-style
// adapters (include/pcl_b200/pcl_b200.h) instead of libpcl.
//
// Same call sequence and parameters as the reference's cloud_cb in SHOT.cpp:298-482 with the
// GeometricConsistency branch (SHOT.cpp:471-483): normals (k) on model and scene → SHOT352 at the
// keypoints (search surface = full cloud) → KdTreeFLANN<SHOT352> nearest-neighbour loop with a
// squared-distance threshold → GeometricConsistencyGrouping::recognize.  ROS, the viewer and ICP are
// outside the hot path and are not part of this harness; clouds come from raw float32 files
// (n x 3, written by the tests) instead of PCD files, keypoints are given (the reference's
// UniformSampling runs on the CPU before the hot path).
//
// usage: shot_recognition <model.f32> <model_kp.f32 | us:leaf> <scene.f32> <scene_kp.f32 | us:leaf> <out_prefix>
//                         [normal_k=10] [descr_rad=0.02] [match_thr=0.25] [cg_size=0.02] [cg_thresh=2] [loop|batch] [gc|hough|hough-shot]
//                         [icp:N] [hv:<normal radius>]
// writes <out_prefix>.corr (b200_corr records), <out_prefix>.T (instances x 16 float),
//        <out_prefix>.inst (int32 count per instance followed by the records)
//        with hough, <out_prefix>.rf (model then scene BOARD frames, 9 float each)
//        with icp:N, <out_prefix>.icp (per instance, first 8: 16 float refined pose, fitness, converged) — the
//        reference's icp_align (SHOT.cpp:177-192) on the model placed by the grouped pose
//        with hv:r (needs icp:N), <out_prefix>.hv (one byte per registered instance) — the reference's hypothesis
//        verification block (SHOT_hypothesis.cpp:631-653) on the ICP-registered instances, with its parameters
//        (:58-64) except the normal radius r (the reference's 5 mm suits its 1 mm scans, not the synthetic clouds)
#include <pcl_b200/pcl_b200.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

namespace pcl = pcl_b200;

typedef pcl::PointXYZRGBA PointType;
typedef pcl::Normal NormalType;
typedef pcl::SHOT352 DescriptorType;

static bool load_cloud(const char *path, pcl::PointCloud<PointType> &cloud) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    std::cerr << "cannot open " << path << std::endl;
    return false;
  }
  float xyz[3];
  while (fread(xyz, sizeof(float), 3, f) == 3) {
    PointType p;
    p.x = xyz[0], p.y = xyz[1], p.z = xyz[2];
    cloud.push_back(p);
  }
  fclose(f);
  return true;
}

int main(int argc, char **argv) {
  if (argc < 6) {
    std::cerr << "usage: " << argv[0] << " model.f32 model_kp.f32 scene.f32 scene_kp.f32 out_prefix [normal_k descr_rad "
              << "match_thr cg_size cg_thresh loop|batch]" << std::endl;
    return 2;
  }
  const int normal_k = argc > 6 ? atoi(argv[6]) : 10;
  const float descr_rad_ = argc > 7 ? (float)atof(argv[7]) : 0.02f;
  const float match_thr = argc > 8 ? (float)atof(argv[8]) : 0.25f;
  const float cg_size_ = argc > 9 ? (float)atof(argv[9]) : 0.02f;
  const float cg_thresh_ = argc > 10 ? (float)atof(argv[10]) : 2.0f;
  const bool batch = argc > 11 && std::string(argv[11]) == "batch";
  const bool loop_single = argc > 11 && std::string(argv[11]) == "loop-single";  // the loop, one device call per query
  const std::string algo = argc > 12 ? argv[12] : "gc";  // the reference's --algorithm Hough|GC
  const bool use_hough = algo == "hough" || algo == "hough-shot";
  const bool board_frames = algo == "hough";
  const float rf_rad_ = 0.02f;  // SHOT.cpp:51
  const int icp_iters = (argc > 13 && std::string(argv[13]).compare(0, 4, "icp:") == 0) ? atoi(argv[13] + 4) : 0;
  const float hv_rad_normals_ = (argc > 14 && std::string(argv[14]).compare(0, 3, "hv:") == 0) ? (float)atof(argv[14] + 3) : 0.f;
  // SHOT_hypothesis.cpp:58-64
  const float hv_clutter_reg_ = 0.001f, hv_inlier_th_ = 0.005f, hv_occlusion_th_ = 0.001f, hv_rad_clutter_ = 0.003f,
              hv_regularizer_ = 0.001f;
  const bool hv_detect_clutter_ = false;

  pcl::PointCloud<PointType>::Ptr model(new pcl::PointCloud<PointType>()), scene(new pcl::PointCloud<PointType>());
  pcl::PointCloud<PointType>::Ptr model_keypoints(new pcl::PointCloud<PointType>()),
      scene_keypoints(new pcl::PointCloud<PointType>());
  pcl::PointCloud<NormalType>::Ptr model_normals(new pcl::PointCloud<NormalType>()),
      scene_normals(new pcl::PointCloud<NormalType>());
  pcl::PointCloud<DescriptorType>::Ptr model_descriptors(new pcl::PointCloud<DescriptorType>()),
      scene_descriptors(new pcl::PointCloud<DescriptorType>());
  if (!load_cloud(argv[1], *model) || !load_cloud(argv[3], *scene)) return 1;
  {
    std::vector<int> indices;  // SHOT.cpp:298-299
    pcl::removeNaNFromPointCloud(*scene, *scene, indices);
  }
  // keypoints: a file, or "us:<leaf>" to extract them like the reference does (SHOT.cpp:314-323)
  for (int side = 0; side < 2; ++side) {
    const std::string spec = argv[side ? 4 : 2];
    pcl::PointCloud<PointType>::Ptr &kp = side ? scene_keypoints : model_keypoints;
    if (spec.rfind("us:", 0) == 0) {
      pcl::UniformSampling<PointType> uniform_sampling;
      uniform_sampling.setInputCloud(side ? scene : model);
      uniform_sampling.setRadiusSearch(atof(spec.c_str() + 3));
      uniform_sampling.filter(*kp);
    } else if (!load_cloud(spec.c_str(), *kp)) {
      return 1;
    }
  }
  std::cout << "Model total points: " << model->size() << "; Selected Keypoints: " << model_keypoints->size() << std::endl;
  std::cout << "Scene total points: " << scene->size() << "; Selected Keypoints: " << scene_keypoints->size() << std::endl;

  //  Compute Normals (one estimator reused for both clouds, as the reference does)
  pcl::NormalEstimationOMP<PointType, NormalType> norm_est;
  norm_est.setKSearch(normal_k);
  norm_est.setInputCloud(model);
  norm_est.compute(*model_normals);
  norm_est.setInputCloud(scene);
  norm_est.compute(*scene_normals);

  //  Compute Descriptor for keypoints
  pcl::SHOTEstimationOMP<PointType, NormalType, DescriptorType> descr_est;
  descr_est.setRadiusSearch(descr_rad_);
  descr_est.setInputCloud(model_keypoints);
  descr_est.setInputNormals(model_normals);
  descr_est.setSearchSurface(model);
  descr_est.compute(*model_descriptors);
  descr_est.setInputCloud(scene_keypoints);
  descr_est.setInputNormals(scene_normals);
  descr_est.setSearchSurface(scene);
  descr_est.compute(*scene_descriptors);
  if (model_descriptors->size() != model_keypoints->size() || scene_descriptors->size() != scene_keypoints->size()) {
    std::cerr << "descriptor computation failed" << std::endl;
    return 1;
  }

  //  Find Model-Scene Correspondences
  pcl::CorrespondencesPtr model_scene_corrs(new pcl::Correspondences());
  const auto t_corr0 = std::chrono::steady_clock::now();
  if (batch) {
    // one call for the whole scene (b200_match): same list as the loop below
    pcl::determineCorrespondences(*model_descriptors, *scene_descriptors, 1, match_thr, *model_scene_corrs);
  } else {
    pcl::KdTreeFLANN<DescriptorType> match_search;
    match_search.setInputCloud(model_descriptors);
    if (loop_single) match_search.setLookAhead(0);
    for (size_t i = 0; i < scene_descriptors->size(); ++i) {
      std::vector<int> neigh_indices(1);
      std::vector<float> neigh_sqr_dists(1);
      if (!std::isfinite(scene_descriptors->at(i).descriptor[0])) continue;  // skipping NaNs
      const int found_neighs = match_search.nearestKSearch(scene_descriptors->at(i), 1, neigh_indices, neigh_sqr_dists);
      if (found_neighs == 1 && neigh_sqr_dists[0] < match_thr)
        model_scene_corrs->push_back(pcl::Correspondence(neigh_indices[0], static_cast<int>(i), neigh_sqr_dists[0]));
    }
  }
  const double corr_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_corr0).count();
  std::cout << "Correspondences found: " << model_scene_corrs->size() << std::endl;
  std::cout << "Correspondence search: " << corr_ms << " ms for " << scene_descriptors->size() << " scene descriptors ("
            << (batch ? "one batched call" : loop_single ? "reference loop, one device call per query"
                                                         : "reference loop, look-ahead batches")
            << "), " << corr_ms * 1e3 / std::max<size_t>(scene_descriptors->size(), 1) << " us per query" << std::endl;

  //  Actual Clustering (GeometricConsistency branch)
  std::vector<pcl::Matrix4f> rototranslations;
  std::vector<pcl::Correspondences> clustered_corrs;
  if (use_hough) {
    // Hough branch (SHOT.cpp:433-470): BOARD frames at the keypoints (SHOT.cpp:441-453), or with
    // "hough-shot" the SHOT frames the descriptor stage produced (pcl::SHOT352::rf).
    pcl::PointCloud<pcl::ReferenceFrame>::Ptr model_rf(new pcl::PointCloud<pcl::ReferenceFrame>()),
        scene_rf(new pcl::PointCloud<pcl::ReferenceFrame>());
    if (board_frames) {
      pcl::BOARDLocalReferenceFrameEstimation<PointType, NormalType, pcl::ReferenceFrame> rf_est;
      rf_est.setFindHoles(true);
      rf_est.setRadiusSearch(rf_rad_);

      rf_est.setInputCloud(model_keypoints);
      rf_est.setInputNormals(model_normals);
      rf_est.setSearchSurface(model);
      rf_est.compute(*model_rf);

      rf_est.setInputCloud(scene_keypoints);
      rf_est.setInputNormals(scene_normals);
      rf_est.setSearchSurface(scene);
      rf_est.compute(*scene_rf);
      FILE *frf = fopen((std::string(argv[5]) + ".rf").c_str(), "wb");
      for (int side = 0; side < 2; ++side)
        for (const pcl::ReferenceFrame &r : (side ? *scene_rf : *model_rf).points) {
          fwrite(r.x_axis, sizeof(float), 3, frf);
          fwrite(r.y_axis, sizeof(float), 3, frf);
          fwrite(r.z_axis, sizeof(float), 3, frf);
        }
      fclose(frf);
    } else {
      for (int side = 0; side < 2; ++side) {
        const pcl::PointCloud<DescriptorType> &d = side ? *scene_descriptors : *model_descriptors;
        pcl::PointCloud<pcl::ReferenceFrame> &rf = side ? *scene_rf : *model_rf;
        rf.resize(d.size());
        for (size_t i = 0; i < d.size(); ++i) {
          memcpy(rf[i].x_axis, d[i].rf + 0, 12);
          memcpy(rf[i].y_axis, d[i].rf + 3, 12);
          memcpy(rf[i].z_axis, d[i].rf + 6, 12);
        }
      }
    }
    pcl::Hough3DGrouping<PointType, PointType, pcl::ReferenceFrame, pcl::ReferenceFrame> clusterer;
    clusterer.setHoughBinSize(cg_size_);
    clusterer.setHoughThreshold(cg_thresh_);
    clusterer.setUseInterpolation(false);
    clusterer.setUseDistanceWeight(true);
    clusterer.setInputCloud(model_keypoints);
    clusterer.setInputRf(model_rf);
    clusterer.setSceneCloud(scene_keypoints);
    clusterer.setSceneRf(scene_rf);
    clusterer.setModelSceneCorrespondences(model_scene_corrs);
    clusterer.recognize(rototranslations, clustered_corrs);
  } else {
    pcl::GeometricConsistencyGrouping<PointType, PointType> gc_clusterer;
    gc_clusterer.setGCSize(cg_size_);
    gc_clusterer.setGCThreshold(cg_thresh_);  // float → int, as in the reference
    gc_clusterer.setInputCloud(model_keypoints);
    gc_clusterer.setSceneCloud(scene_keypoints);
    gc_clusterer.setModelSceneCorrespondences(model_scene_corrs);
    gc_clusterer.recognize(rototranslations, clustered_corrs);
  }

  std::cout << "Model instances found: " << rototranslations.size() << std::endl;
  for (size_t i = 0; i < rototranslations.size() && i < 3; ++i) {
    std::cout << "\n    Instance " << i + 1 << ":" << std::endl;
    std::cout << "        Correspondences belonging to this instance: " << clustered_corrs[i].size() << std::endl;
    const pcl::Matrix4f &T = rototranslations[i];
    printf("            | %6.3f %6.3f %6.3f | \n", T(0, 0), T(0, 1), T(0, 2));
    printf("        R = | %6.3f %6.3f %6.3f | \n", T(1, 0), T(1, 1), T(1, 2));
    printf("            | %6.3f %6.3f %6.3f | \n", T(2, 0), T(2, 1), T(2, 2));
    printf("        t = < %0.3f, %0.3f, %0.3f >\n", T(0, 3), T(1, 3), T(2, 3));
  }

  const std::string prefix = argv[5];
  FILE *f = fopen((prefix + ".corr").c_str(), "wb");
  if (!model_scene_corrs->empty()) fwrite(model_scene_corrs->data(), sizeof(pcl::Correspondence), model_scene_corrs->size(), f);
  fclose(f);
  f = fopen((prefix + ".T").c_str(), "wb");
  for (size_t i = 0; i < rototranslations.size(); ++i) fwrite(rototranslations[i].m, sizeof(float), 16, f);
  fclose(f);
  f = fopen((prefix + ".inst").c_str(), "wb");
  for (size_t i = 0; i < clustered_corrs.size(); ++i) {
    const int32_t n = (int32_t)clustered_corrs[i].size();
    fwrite(&n, sizeof(n), 1, f);
    if (n) fwrite(clustered_corrs[i].data(), sizeof(pcl::Correspondence), (size_t)n, f);
  }
  fclose(f);
  std::vector<pcl::PointCloud<PointType>::ConstPtr> registered_instances;
  if (icp_iters > 0) {
    f = fopen((prefix + ".icp").c_str(), "wb");
    for (size_t i = 0; i < rototranslations.size() && i < 8; ++i) {
      // SHOT.cpp: transformPointCloud(*model, *rotated_model, rototranslations[i]); icp_align(scene, rotated_model)
      pcl::PointCloud<PointType>::Ptr rotated_model(new pcl::PointCloud<PointType>());
      pcl::transformPointCloud(*model, *rotated_model, rototranslations[i]);
      pcl::IterativeClosestPoint<PointType, PointType> icp;
      icp.setMaximumIterations(icp_iters);
      icp.setInputSource(rotated_model);
      icp.setInputTarget(scene);
      pcl::PointCloud<PointType> cloud_icp;
      icp.align(cloud_icp);
      const double score = icp.getFitnessScore();
      if (i < 3) printf("\nICP has converged, score is %+.0e\n", score);
      float rec[18];
      memcpy(rec, icp.getFinalTransformation().m, sizeof(float) * 16);
      rec[16] = (float)score;
      rec[17] = icp.hasConverged() ? 1.f : 0.f;
      fwrite(rec, sizeof(float), 18, f);
      registered_instances.push_back(pcl::PointCloud<PointType>::ConstPtr(new pcl::PointCloud<PointType>(cloud_icp)));
    }
    fclose(f);
  }
  if (hv_rad_normals_ > 0.f && registered_instances.size() > 0) {
    // SHOT_hypothesis.cpp:631-653, statement for statement
    std::cout << "--- Hypotheses Verification ---" << std::endl;
    std::vector<bool> hypotheses_mask;  // Mask Vector to identify positive hypotheses
    pcl::GlobalHypothesesVerification<PointType, PointType> GoHv;
    GoHv.setSceneCloud(scene);                      // Scene Cloud
    GoHv.addModels(registered_instances, true);     // Models to verify
    GoHv.setInlierThreshold(hv_inlier_th_);
    GoHv.setOcclusionThreshold(hv_occlusion_th_);
    GoHv.setRegularizer(hv_regularizer_);
    GoHv.setRadiusClutter(hv_rad_clutter_);
    GoHv.setClutterRegularizer(hv_clutter_reg_);
    GoHv.setDetectClutter(hv_detect_clutter_);
    GoHv.setRadiusNormals(hv_rad_normals_);
    GoHv.verify();
    GoHv.getMask(hypotheses_mask);  // i-element TRUE if hvModels[i] verifies hypotheses
    f = fopen((prefix + ".hv").c_str(), "wb");
    for (size_t i = 0; i < hypotheses_mask.size(); i++) {
      if (hypotheses_mask[i]) std::cout << "Instance " << i << " is GOOD! <---" << std::endl;
      const unsigned char b = hypotheses_mask[i] ? 1 : 0;
      fwrite(&b, 1, 1, f);
    }
    fclose(f);
  }
  return 0;
}
