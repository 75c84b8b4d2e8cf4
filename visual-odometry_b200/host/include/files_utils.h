// files_utils.h — declarations of the reference's text I/O helpers (include/files_utils.h).
// File parsing is outside the GPU hot path: the functions themselves are the reference's own
// src/files_utils.cpp, compiled unchanged by build_dropin.sh.  This header only exists so that
// the translation unit sees THIS repo's defs.h / PointCloud.h instead of the reference's.
#pragma once
#include <dirent.h>

#include <fstream>
#include <iostream>
#include <regex>
#include <set>
#include <sstream>
#include <string>
#include <unordered_set>

#include "PointCloud.h"
#include "defs.h"

// one vector per line, coefficients separated by blanks (Eigen's default matrix format)
template <typename vec_type>
void write_eigen_vectors_to_file(const std::string& file_path,
                                 const std::vector<vec_type, Eigen::aligned_allocator<vec_type>>& vectors) {
  std::ofstream out(file_path);
  if (!out.is_open()) {
    std::cout << "Error opening file" << std::endl;
    return;
  }
  for (size_t i = 0; i < vectors.size(); ++i) out << vectors[i].transpose() << std::endl;
}

bool get_file_names(const std::string& path, std::set<std::string>& files, const std::regex& pattern);
bool get_meas_content(const std::string& file_path, Vector10fVector& appearances,
                      Vector3fVector& features, const bool& is_world = false);
bool get_meas_content(const std::string& file_path, PointCloudVector<2>& points);
bool get_camera_params(const std::string& file_path, std::vector<int>& int_params,
                       Eigen::Matrix3f& k, Eigen::Isometry3f& H);
void save_trajectory(const std::string& file_path, const IsometryVector& vector,
                     const Eigen::Isometry3f& cameraInRobot = Eigen::Isometry3f::Identity(),
                     const bool& save_rotation = false);
bool save_gt_trajectory(const std::string& file_path);
