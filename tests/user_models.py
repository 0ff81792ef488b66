"""CUDA sources of user models for the run-time compiled path (ilqr_register_model): what a user of the reference's
`class Model` plugin surface (include/model.h:6-21) writes to run a model of their own on the GPU."""

# The built-in DoubleIntegrator twin (ilqr_b200/csrc/models.cuh, reference include/double_integrator.h) written out
# again as a user would: the same expressions, so the run-time compiled kernel must agree with the built-in one bit
# for bit.  model_params[0..3] = goal.
DOUBLE_INTEGRATOR_CLONE = r"""
struct UserDoubleIntegrator {
  static constexpr int N = 4;
  static constexpr int M = 2;
  template <typename S>
  ILQR_HD static void dynamics(const S *x, const S *u, const S *, S *dx) {
    const S mass = 1;
    dx[0] = x[2];
    dx[1] = x[3];
    dx[2] = u[0] / mass;
    dx[3] = u[1] / mass;
  }
  static constexpr unsigned kConfigVars = 0;
  template <typename S>
  struct Config {};
  template <typename S>
  ILQR_HD static void configure(const S *, const S *, Config<S> &) {}
  template <typename S>
  ILQR_HD static void dynamics_cfg(const Config<S> &, const S *x, const S *u, const S *mp, S *dx) { dynamics(x, u, mp, dx); }
  template <typename S>
  ILQR_HD static S quad(const S *x, const S *mp, S scale) {
    const S hx[4] = {S(1), S(1), S(0.2), S(0.2)};
    S acc = 0;
    for (int i = 0; i < 4; i++) {
      const S e = mp[i] - x[i];
      acc += (e * (scale * hx[i])) * e;
    }
    return acc;
  }
  template <typename S>
  ILQR_HD static S cost(const S *x, const S *u, const S *mp) { return quad(x, mp, S(1)) + (u[0] * u[0] + u[1] * u[1]); }
  template <typename S>
  ILQR_HD static S final_cost(const S *x, const S *mp) { return quad(x, mp, S(10)); }
  template <typename S>
  ILQR_HD static S cost_d1(int c, const S *x, const S *u, const S *mp, bool terminal) {
    const S sc = terminal ? S(10) : S(1);
    if (c < N) {
      const S h = c < 2 ? S(1) : S(0.2);
      return S(-2.0) * (sc * h) * (mp[c] - x[c]);
    }
    return terminal ? S(0) : 2 * u[c - N];
  }
  template <typename S>
  ILQR_HD static S cost_d2(int c, int d, const S *, const S *, const S *, bool terminal) {
    if (c != d) return S(0);
    if (c >= N) return S(2);
    const S sc = terminal ? S(10) : S(1);
    const S h = c < 2 ? S(1) : S(0.2);
    return S(2.0) * (sc * h);
  }
};
"""

# A model the library does not ship: a damped pendulum, n = 2 (angle, rate), m = 1 (torque), swing-up to
# model_params[0] (the goal angle).  The sine of the angle is the configuration-dependent part, shared by the
# finite-difference points that perturb the rate or the torque.
PENDULUM = r"""
struct Pendulum {
  static constexpr int N = 2;
  static constexpr int M = 1;
  static constexpr unsigned kConfigVars = 0x1;
  template <typename S>
  struct Config {
    S sin_q;
  };
  template <typename S>
  ILQR_HD static void configure(const S *x, const S *, Config<S> &cf) {
    S s, c;
    ilqr::sincos_det(x[0], &s, &c);
    cf.sin_q = s;
  }
  template <typename S>
  ILQR_HD static void dynamics_cfg(const Config<S> &cf, const S *x, const S *u, const S *, S *dx) {
    const S g = S(9.81), l = S(1), mass = S(1), damping = S(0.1);
    dx[0] = x[1];
    dx[1] = (u[0] - damping * x[1] - mass * g * l * cf.sin_q) / (mass * l * l);
  }
  template <typename S>
  ILQR_HD static void dynamics(const S *x, const S *u, const S *mp, S *dx) {
    Config<S> cf;
    configure(x, mp, cf);
    dynamics_cfg(cf, x, u, mp, dx);
  }
  template <typename S>
  ILQR_HD static S cost(const S *x, const S *u, const S *mp) {
    const S e = mp[0] - x[0];
    return S(0.01) * (e * e) + S(0.001) * (x[1] * x[1]) + S(0.05) * (u[0] * u[0]);
  }
  template <typename S>
  ILQR_HD static S final_cost(const S *x, const S *mp) {
    const S e = mp[0] - x[0];
    return S(100) * (e * e) + S(10) * (x[1] * x[1]);
  }
  template <typename S>
  ILQR_HD static S cost_d1(int c, const S *x, const S *u, const S *mp, bool terminal) {
    const S e = mp[0] - x[0];
    if (c == 0) return (terminal ? S(-200) : S(-0.02)) * e;
    if (c == 1) return (terminal ? S(20) : S(0.002)) * x[1];
    return terminal ? S(0) : S(0.1) * u[0];
  }
  template <typename S>
  ILQR_HD static S cost_d2(int c, int d, const S *, const S *, const S *, bool terminal) {
    if (c != d) return S(0);
    if (c == 0) return terminal ? S(200) : S(0.02);
    if (c == 1) return terminal ? S(20) : S(0.002);
    return S(0.1);
  }
};
"""

BROKEN = "struct Broken { static constexpr int N = 4; static constexpr int M = 1; };"
