#pragma once
#include <limits>
#include "../point_cloud.h"
namespace pcl {
// pcl::Registration as far as a subclass sees it (PCL 1.8 registration.h): align() copies the source and calls the virtual
// computeTransformation; the protected members are the ones the reference's NDT classes write.
template <typename PointSource, typename PointTarget>
class Registration {
 public:
  typedef boost::shared_ptr<Registration<PointSource, PointTarget> > Ptr;
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  virtual ~Registration() {}
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  Eigen::Matrix4f getFinalTransformation() { return final_transformation_; }
  bool hasConverged() { return converged_; }
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) { return max_range; }
  void align(PointCloudSource& output, const Eigen::Matrix4f& guess) { output = *input_; computeTransformation(output, guess); }

 protected:
  std::string reg_name_;
  int nr_iterations_ = 0, max_iterations_ = 10;
  double transformation_epsilon_ = 0.0;
  bool converged_ = false;
  Eigen::Matrix4f final_transformation_;
  PointCloudTargetConstPtr target_;
  PointCloudSourceConstPtr input_;
  virtual void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) = 0;
};
}  // namespace pcl
