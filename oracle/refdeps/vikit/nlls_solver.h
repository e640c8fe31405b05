// Stand-in for rpg_vikit vikit_common/nlls_solver.h + nlls_solver_impl.hpp — TEST INFRASTRUCTURE.
//
// vk::NLLSSolver<D,T> is the base class of plsvo::SparseImgAlign; its Gauss-Newton driver is
// part of the hot path but lives in the un-vendored dependency, so it is restated here from
// the published source: optional weight-scale pre-pass, per iteration { startIteration; zero
// H/Jres; n_meas_ = 0; computeResiduals(linearize); solve(); stop on failure or chi2 increase
// with rollback; update(); finishIteration(); stop when norm_max(x) <= eps }.
// Levenberg-Marquardt is never selected on this path (frame_handler_mono.cpp:272-273).
#ifndef PLSVO_REFDEPS_VIKIT_NLLS_SOLVER
#define PLSVO_REFDEPS_VIKIT_NLLS_SOLVER
#include <Eigen/Core>
#include <vikit/math_utils.h>
#include <vikit/robust_cost.h>
#include <iostream>
#include <stdexcept>

namespace vk {
using namespace std;
using namespace Eigen;

template <int D, typename T>
class NLLSSolver {
 public:
  typedef T ModelType;
  enum Method { GaussNewton, LevenbergMarquardt };
  enum ScaleEstimatorType { UnitScale, TDistScale, MADScale, NormalScale };
  enum WeightFunctionType { UnitWeight, TDistWeight, TukeyWeight, HuberWeight };

 protected:
  Matrix<double, D, D> H_;     //!< Hessian approximation
  Matrix<double, D, 1> Jres_;  //!< Jacobian x Residual
  Matrix<double, D, 1> x_;     //!< update step
  bool have_prior_;
  ModelType prior_;
  Matrix<double, D, D> I_prior_;
  double chi2_;
  double rho_;
  Method method_;

  virtual double computeResiduals(const ModelType& model, bool linearize_system, bool compute_weight_scale) = 0;
  virtual int solve() = 0;
  virtual void update(const ModelType& old_model, ModelType& new_model) = 0;
  virtual void applyPrior(const ModelType&) {}
  virtual void startIteration() {}
  virtual void finishIteration() {}

 public:
  double mu_init_, mu_;
  double nu_init_, nu_;
  size_t n_iter_init_, n_iter_;
  size_t n_trials_;
  size_t n_trials_max_;
  size_t n_meas_;
  bool stop_;
  bool verbose_;
  double eps_;
  size_t iter_;

  bool use_weights_;
  float scale_;
  robust_cost::ScaleEstimatorPtr scale_estimator_;
  robust_cost::WeightFunctionPtr weight_function_;

  NLLSSolver()
      : have_prior_(false), method_(LevenbergMarquardt), mu_init_(0.01f), mu_(mu_init_), nu_init_(2.0),
        nu_(nu_init_), n_iter_init_(15), n_iter_(n_iter_init_), n_trials_(0), n_trials_max_(5), n_meas_(0),
        stop_(false), verbose_(true), eps_(0.0000000001), iter_(0), use_weights_(false), scale_(0.0),
        scale_estimator_(), weight_function_() {}
  virtual ~NLLSSolver() {}

  void optimize(ModelType& model) {
    if (method_ == GaussNewton)
      optimizeGaussNewton(model);
    else
      throw std::runtime_error("vikit stand-in: Levenberg-Marquardt is not on the hot path");
  }

  void optimizeGaussNewton(ModelType& model) {
    // Compute weight scale
    if (use_weights_) computeResiduals(model, false, true);

    // Save the old model to rollback in case of unsuccessful update
    ModelType old_model(model);

    for (iter_ = 0; iter_ < n_iter_; ++iter_) {
      rho_ = 0;
      startIteration();

      H_.setZero();
      Jres_.setZero();

      // compute initial error
      n_meas_ = 0;
      double new_chi2 = computeResiduals(model, true, false);

      if (have_prior_) applyPrior(model);

      // solve the linear system
      if (!solve()) {
        // matrix was singular and could not be computed
        if (verbose_) std::cout << "Matrix is close to singular! Stop Optimizing." << std::endl;
        stop_ = true;
      }

      // check if error increased since last optimization
      if ((iter_ > 0 && new_chi2 > chi2_) || stop_) {
        if (verbose_)
          std::cout << "It. " << iter_ << "\t Failure \t new_chi2 = " << new_chi2 << "\t Error increased. Stop optimizing."
                    << std::endl;
        model = old_model;  // rollback
        break;
      }

      // update the model
      ModelType new_model;
      update(model, new_model);
      old_model = model;
      model = new_model;

      chi2_ = new_chi2;

      if (verbose_)
        std::cout << "It. " << iter_ << "\t Success \t new_chi2 = " << new_chi2 << "\t n_meas = " << n_meas_
                  << "\t x_norm = " << vk::norm_max(x_) << std::endl;

      finishIteration();

      // stop when converged, i.e. update step too small
      if (vk::norm_max(x_) <= eps_) break;
    }
  }

  void reset() {
    have_prior_ = false;
    chi2_ = 1e10;
    mu_ = mu_init_;
    nu_ = nu_init_;
    n_meas_ = 0;
    n_iter_ = n_iter_init_;
    iter_ = 0;
    stop_ = false;
  }

  const double& getChi2() const { return chi2_; }
  const Matrix<double, D, D>& getInformationMatrix() const { return H_; }
  void setRobustCostFunction(ScaleEstimatorType, WeightFunctionType) {}
  void setPrior(const T& prior, const Matrix<double, D, D>& Information) {
    have_prior_ = true;
    prior_ = prior;
    I_prior_ = Information;
  }
};

}  // namespace vk
#endif
