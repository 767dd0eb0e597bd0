// fpfh_recognition.cpp — the FPFH variant of the reference's recognition callback
// (FPFH_demo.cpp:405-538) through the PCL-style adapters: normals on the keypoint clouds by radius
// (FPFH_demo.cpp:416-420, 486-492), FPFHEstimation with input = surface = keypoints
// (FPFH_demo.cpp:422-428, 505-510), KdTreeFLANN<FPFHSignature33> k = 2 with the ratio test
// (FPFH_demo.cpp:516-538), then GeometricConsistencyGrouping (the north-star grouping; the reference's
// default there is Hough3D, which is a "next" row of SURVEY.md §8(f)).
//
// usage: fpfh_recognition <model_kp.f32> <scene_kp.f32> <out_prefix> [radius=0.05] [cg_size=0.02] [cg_thresh=2]
#include <pcl_b200/pcl_b200.h>

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

namespace pcl = pcl_b200;

typedef pcl::PointXYZRGBA PointType;
typedef pcl::Normal NormalType;
typedef pcl::FPFHSignature33 DescriptorType;

static bool load_cloud(const char *path, pcl::PointCloud<PointType> &cloud) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
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
  if (argc < 4) {
    std::cerr << "usage: " << argv[0] << " model_kp.f32 scene_kp.f32 out_prefix [radius cg_size cg_thresh]" << std::endl;
    return 2;
  }
  const double radius = argc > 4 ? atof(argv[4]) : 0.05;
  const float cg_size_ = argc > 5 ? (float)atof(argv[5]) : 0.02f;
  const float cg_thresh_ = argc > 6 ? (float)atof(argv[6]) : 2.0f;
  pcl::PointCloud<PointType>::Ptr model_keypoints(new pcl::PointCloud<PointType>()),
      scene_keypoints(new pcl::PointCloud<PointType>());
  pcl::PointCloud<NormalType>::Ptr model_normals(new pcl::PointCloud<NormalType>()),
      scene_normals(new pcl::PointCloud<NormalType>());
  pcl::PointCloud<DescriptorType>::Ptr model_descriptors(new pcl::PointCloud<DescriptorType>()),
      scene_descriptors(new pcl::PointCloud<DescriptorType>());
  if (!load_cloud(argv[1], *model_keypoints) || !load_cloud(argv[2], *scene_keypoints)) return 1;

  pcl::search::KdTree<PointType>::Ptr kdtree(new pcl::search::KdTree<PointType>());
  pcl::NormalEstimationOMP<PointType, NormalType> norm_est;
  norm_est.setSearchMethod(kdtree);
  norm_est.setRadiusSearch(radius);
  norm_est.setInputCloud(scene_keypoints);
  norm_est.compute(*scene_normals);
  norm_est.setInputCloud(model_keypoints);
  norm_est.compute(*model_normals);

  pcl::FPFHEstimation<PointType, NormalType, DescriptorType> fpfh;
  fpfh.setSearchMethod(kdtree);
  fpfh.setRadiusSearch(radius);
  fpfh.setInputCloud(scene_keypoints);
  fpfh.setInputNormals(scene_normals);
  fpfh.compute(*scene_descriptors);
  fpfh.setInputCloud(model_keypoints);
  fpfh.setInputNormals(model_normals);
  fpfh.compute(*model_descriptors);

  pcl::CorrespondencesPtr model_scene_corrs(new pcl::Correspondences());
  pcl::KdTreeFLANN<DescriptorType> match_search;
  match_search.setInputCloud(model_descriptors);
  for (size_t i = 0; i < scene_descriptors->size(); ++i) {
    std::vector<int> neigh_indices(2);
    std::vector<float> neigh_sqr_dists(2);
    if (!std::isfinite(scene_descriptors->at(i).histogram[0])) continue;
    const int found_neighs = match_search.nearestKSearch(scene_descriptors->at(i), 2, neigh_indices, neigh_sqr_dists);
    if (found_neighs < 1) continue;
    const double tau = found_neighs > 1 ? (double)neigh_sqr_dists[0] / neigh_sqr_dists[1] : 0.0;
    if (tau <= 1) model_scene_corrs->push_back(pcl::Correspondence(neigh_indices[0], static_cast<int>(i), neigh_sqr_dists[0]));
  }
  std::cout << "Correspondences found: " << model_scene_corrs->size() << std::endl;

  std::vector<pcl::Matrix4f> rototranslations;
  std::vector<pcl::Correspondences> clustered_corrs;
  pcl::GeometricConsistencyGrouping<PointType, PointType> gc_clusterer;
  gc_clusterer.setGCSize(cg_size_);
  gc_clusterer.setGCThreshold(cg_thresh_);
  gc_clusterer.setInputCloud(model_keypoints);
  gc_clusterer.setSceneCloud(scene_keypoints);
  gc_clusterer.setModelSceneCorrespondences(model_scene_corrs);
  gc_clusterer.recognize(rototranslations, clustered_corrs);
  std::cout << "Model instances found: " << rototranslations.size() << std::endl;

  const std::string prefix = argv[3];
  FILE *f = fopen((prefix + ".desc").c_str(), "wb");
  for (size_t i = 0; i < scene_descriptors->size(); ++i) fwrite(scene_descriptors->at(i).histogram, sizeof(float), 33, f);
  fclose(f);
  // the normals the descriptors were computed from (stage-by-stage parity checks start from identical inputs)
  f = fopen((prefix + ".normals").c_str(), "wb");
  for (size_t i = 0; i < scene_normals->size(); ++i) {
    const float n4[4] = {scene_normals->at(i).normal_x, scene_normals->at(i).normal_y, scene_normals->at(i).normal_z,
                         scene_normals->at(i).curvature};
    fwrite(n4, sizeof(float), 4, f);
  }
  fclose(f);
  f = fopen((prefix + ".corr").c_str(), "wb");
  if (!model_scene_corrs->empty()) fwrite(model_scene_corrs->data(), sizeof(pcl::Correspondence), model_scene_corrs->size(), f);
  fclose(f);
  f = fopen((prefix + ".T").c_str(), "wb");
  for (size_t i = 0; i < rototranslations.size(); ++i) fwrite(rototranslations[i].m, sizeof(float), 16, f);
  fclose(f);
  return 0;
}
