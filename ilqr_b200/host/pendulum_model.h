/*
 * pendulum_model.h — a `Model` subclass of the USER's own, written against the reference's plugin surface
 * (include/model.h:6-21: dynamics, cost, final_cost, u_min / u_max, x_dims / u_dims): a damped pendulum, n = 2 (angle,
 * rate), m = 1 (torque), swing-up to `goal`.  Used by user_model_demo.cpp (where it runs on the GPU through its
 * registered device twin, tests/user_models.py: PENDULUM) and by oracle/ref_harness.cpp (where the UNMODIFIED
 * reference's own `iLQR` class solves it on the CPU: the oracle-side check of the user-model path).
 * Include after the reference's common.h / model.h.
 */
#ifndef ILQR_PENDULUM_MODEL_H_
#define ILQR_PENDULUM_MODEL_H_

class Pendulum : public Model {
 public:
  explicit Pendulum(double goal_angle = 3.141592653589793) : goal(goal_angle) {
    x_dims = 2;
    u_dims = 1;
    u_min.resize(1);
    u_max.resize(1);
    u_min << -2.0;
    u_max << 2.0;
  }
  double goal;
  virtual VectorXd dynamics(const VectorXd &x, const VectorXd &u) {
    const double g = 9.81, l = 1, mass = 1, damping = 0.1;
    VectorXd dx(2);
    dx(0) = x(1);
    dx(1) = (u(0) - damping * x(1) - mass * g * l * sin(x(0))) / (mass * l * l);
    return dx;
  }
  virtual double cost(const VectorXd &x, const VectorXd &u) {
    const double e = goal - x(0);
    return 0.01 * (e * e) + 0.001 * (x(1) * x(1)) + 0.05 * (u(0) * u(0));
  }
  virtual double final_cost(const VectorXd &x) {
    const double e = goal - x(0);
    return 100 * (e * e) + 10 * (x(1) * x(1));
  }
};

#endif
