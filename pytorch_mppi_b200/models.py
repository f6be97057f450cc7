"""Registry of analytic models whose dynamics + running cost are compiled into the fused kernel.

The reference's plugin surface is a pair of Python callables ``dynamics(state, action)`` /
``running_cost(state, action)`` (mppi.py:63-64, README.md:53-58).  Arbitrary callables cannot be
inlined into CUDA, so the fused path recognises callables that are *bound methods of a registered
model object*: pass ``model.dynamics`` / ``model.running_cost`` (and optionally
``model.terminal_cost``) to the controller exactly as you would pass your own functions, and the
controller launches the fused sm_100a kernel.  The same bound methods are ordinary torch functions
(any device, any dtype), so they also serve as the simulator you step your real system with and as
the callables of the per-step path.

Anything else (a torch MLP, a lambda, a step-dependent function) takes the per-step path:
the T-loop stays in Python with sampling / accumulation / softmin in CUDA kernels.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import _cabi


class AnalyticModel:
    """Base class: a model known to the CUDA registry (include/mppi_b200.h `MppiModel`)."""
    model_id: int = 0
    nx: int = 0
    nu: int = 0

    def param_blob(self) -> list:
        """The `model_params` doubles of MppiFusedParams for this instance."""
        raise NotImplementedError

    def param_blob_ext(self) -> list:
        """Extra parameters too large for `model_params` (e.g. network weights); [] if none."""
        return []

    @property
    def has_terminal(self) -> bool:
        return False

    # -- recognition of bound methods --------------------------------------------------------
    @staticmethod
    def owner_of(fn) -> Optional["AnalyticModel"]:
        owner = getattr(fn, "__self__", None)
        return owner if isinstance(owner, AnalyticModel) else None


class Pendulum(AnalyticModel):
    """Gym pendulum swing-up, the model of BASELINE configs 1/2/5
    (/root/reference/tests/pendulum.py:30-60): state (theta, theta_dot), one torque input.

        u      <- clamp(u, +-max_torque)
        thdot' =  clip(thdot + (3g/(2l) sin(theta) + 3/(m l^2) u) dt, +-max_speed)
        theta' =  theta + thdot' dt
        cost   =  angle_normalize(theta')^2 + w_thdot thdot'^2
    """
    model_id = _cabi.MODEL_PENDULUM
    nx, nu = 2, 1

    def __init__(self, g=10.0, m=1.0, l=1.0, dt=0.05, max_torque=2.0, max_speed=8.0, w_thdot=0.1):
        self.g, self.m, self.l, self.dt = float(g), float(m), float(l), float(dt)
        self.max_torque, self.max_speed, self.w_thdot = float(max_torque), float(max_speed), float(w_thdot)

    def param_blob(self):
        return [self.g, self.m, self.l, self.dt, self.max_torque, self.max_speed, self.w_thdot]

    def dynamics(self, state, action):
        th = state[:, 0:1]
        thdot = state[:, 1:2]
        u = torch.clamp(action[:, 0:1], -self.max_torque, self.max_torque)
        acc = (3 * self.g / (2 * self.l)) * torch.sin(th) + (3.0 / (self.m * self.l ** 2)) * u
        new_thdot = torch.clamp(thdot + acc * self.dt, -self.max_speed, self.max_speed)
        new_th = th + new_thdot * self.dt
        return torch.cat((new_th, new_thdot), dim=1)

    @staticmethod
    def angle_normalize(x):
        return ((x + math.pi) % (2 * math.pi)) - math.pi

    def running_cost(self, state, action):
        return self.angle_normalize(state[:, 0]) ** 2 + self.w_thdot * state[:, 1] ** 2


class LinearPoint(AnalyticModel):
    """2-D point mass with linear-delta dynamics ``x + u @ B.T``, a quadratic goal cost
    ``(g-x)^T Q (g-x)``, optional action cost ``u^T R u``, up to three Gaussian hill costs
    ``h exp(-(c-x)^T Q_h (c-x))`` and an optional terminal cost ``terminal_scale * state_cost(x_T)``.

    `LinearPoint.unit_test_env()` is the fixture of the reference's test-suite
    (/root/reference/tests/test_mppi.py:24-51); `LinearPoint.toy2d_nav()` is the Toy2D navigation
    environment of /root/reference/tests/smooth_mppi.py:79-142 (BASELINE config 3).
    """
    model_id = _cabi.MODEL_LINEAR_POINT
    nx, nu = 2, 2
    MAX_HILLS = 3

    def __init__(self, B, goal, Q=None, R=None, hills: Sequence = (), terminal_scale: float = 0.0):
        self.B = [[float(v) for v in row] for row in B]
        self.goal = [float(v) for v in goal]
        self.Q = [[1.0, 0.0], [0.0, 1.0]] if Q is None else [[float(v) for v in row] for row in Q]
        self.R = None if R is None else [[float(v) for v in row] for row in R]
        self.hills = [([[float(v) for v in row] for row in q], [float(v) for v in c], float(h)) for q, c, h in hills]
        if len(self.hills) > self.MAX_HILLS:
            raise ValueError(f"at most {self.MAX_HILLS} hills are compiled into the fused kernel")
        self.terminal_scale = float(terminal_scale)
        self._cache = {}

    @classmethod
    def unit_test_env(cls, terminal_scale=0.0):
        return cls(B=[[1.0, 0.0], [0.0, -1.0]], goal=[2.0, 2.0], terminal_scale=terminal_scale)

    @classmethod
    def toy2d_nav(cls, terminal_scale=10.0, r=0.01):
        return cls(B=[[0.5, 0.0], [0.0, -0.5]], goal=[2.0, 2.0], R=[[r, 0.0], [0.0, r]],
                   hills=[([[0.25, 0.125], [0.125, 0.25]], [-0.5, -1.0], 200.0)], terminal_scale=terminal_scale)

    @property
    def has_terminal(self):
        return self.terminal_scale != 0.0

    def param_blob(self):
        flat = lambda m: [m[0][0], m[0][1], m[1][0], m[1][1]]
        b = flat(self.B) + self.goal + flat(self.Q)
        b += [1.0 if self.R is not None else 0.0] + (flat(self.R) if self.R is not None else [0.0] * 4)
        b += [self.terminal_scale, float(len(self.hills))]
        for q, c, h in self.hills:
            b += flat(q) + c + [h]
        b += [0.0] * (7 * (self.MAX_HILLS - len(self.hills)))
        return b

    # ---- torch implementations (any device/dtype) ---------------------------------------------
    def _t(self, name, like):
        key = (name, like.device, like.dtype)
        if key not in self._cache:
            src = {"B": self.B, "goal": self.goal, "Q": self.Q, "R": self.R}.get(name)
            if src is None and name.startswith("hq"):
                src = self.hills[int(name[2:])][0]
            if src is None and name.startswith("hc"):
                src = self.hills[int(name[2:])][1]
            self._cache[key] = torch.tensor(src, device=like.device, dtype=like.dtype)
        return self._cache[key]

    @staticmethod
    def _quad(d, Q):
        return (d * (d @ Q.transpose(0, 1))).sum(dim=-1)

    def dynamics(self, state, action):
        return state + action @ self._t("B", state).transpose(0, 1)

    def state_cost(self, state):
        c = self._quad(self._t("goal", state) - state, self._t("Q", state))
        for i, (_, _, h) in enumerate(self.hills):
            c = c + h * torch.exp(-self._quad(self._t(f"hc{i}", state) - state, self._t(f"hq{i}", state)))
        return c

    def running_cost(self, state, action):
        c = self.state_cost(state)
        if self.R is not None:
            c = c + self._quad(action, self._t("R", state))
        return c

    def terminal_cost(self, states, actions):
        return self.terminal_scale * self.state_cost(states[..., -1, :])


class PendulumMLP(AnalyticModel):
    """Learned pendulum dynamics of BASELINE config 4 (/root/reference/tests/pendulum_approximate.py:47-67):
    ``x' = x + net([x, clamp(u)])`` with ``net = Linear(3,32)-Tanh-Linear(32,32)-Tanh-Linear(32,2)``, then the
    angle is wrapped to [-pi, pi); running cost as the analytic pendulum.

    Pass ``model.dynamics`` / ``model.running_cost`` to the controller: the fused kernel then evaluates the
    network inside the rollout (weights in the kernel's constant parameter block).  The torch callables
    below evaluate the same ``net`` module, so they are the stepped-route / simulator definition too.
    Call ``refresh()`` after retraining ``net`` (the controller re-reads the weights when repacked).
    """
    model_id = _cabi.MODEL_PENDULUM_MLP
    nx, nu = 2, 1
    H = 32

    def __init__(self, net: torch.nn.Sequential, max_torque=2.0, w_thdot=0.1, fast_tanh=False, tensor_cores="auto"):
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        act = [m for m in net if not isinstance(m, torch.nn.Linear)]
        shapes = [tuple(l.weight.shape) for l in lin]
        if shapes != [(self.H, 3), (self.H, self.H), (2, self.H)] or not all(isinstance(a, torch.nn.Tanh) for a in act):
            raise ValueError("PendulumMLP expects Linear(3,32)-Tanh-Linear(32,32)-Tanh-Linear(32,2); got " + str(shapes))
        self.net = net
        self.max_torque, self.w_thdot, self.fast_tanh = float(max_torque), float(w_thdot), bool(fast_tanh)
        # tensor_cores (fp32 controllers only): the three layers run as tcgen05 MMAs with TMEM accumulators
        # (csrc/mppi_mlp_tc.cuh).  True / "bf16x3": operands are hi/lo-split bf16, layer outputs agree with the
        # fp32 FFMA kernel to ~1e-5 relative.  "bf16": plain bf16 operands for the two hidden-layer products
        # (~2^-8 relative; the state inputs stay split), the fastest route.
        # "auto" (default): "bf16x3" wherever the tensor-core kernel exists (fp32 controllers, one environment), the FFMA
        # kernel elsewhere (fp64, MPPI_Batched); False pins the FFMA kernel.
        if not any(tensor_cores is v or tensor_cores == v for v in (False, True, None, "auto", "bf16x3", "bf16")):
            raise ValueError("tensor_cores must be 'auto', False, True, 'bf16x3' or 'bf16'")
        if tensor_cores is True:
            tensor_cores = "bf16x3"
        self.tensor_cores = {False: 0, None: 0, "bf16x3": 1, "bf16": 2, "auto": 3}[tensor_cores]

    def param_blob(self):
        return [self.max_torque, self.w_thdot, 1.0 if self.fast_tanh else 0.0, float(self.tensor_cores)]

    def param_blob_ext(self):
        out = []
        with torch.no_grad():
            for m in self.net:
                if isinstance(m, torch.nn.Linear):
                    out += m.weight.detach().double().cpu().reshape(-1).tolist()
                    out += m.bias.detach().double().cpu().reshape(-1).tolist()
        return out

    def dynamics(self, state, action):
        with torch.no_grad():
            u = torch.clamp(action[:, 0:1], -self.max_torque, self.max_torque)
            nxt = state + self.net(torch.cat((state, u), dim=1))
            th = Pendulum.angle_normalize(nxt[:, 0])
            return torch.stack((th, nxt[:, 1]), dim=1)

    def running_cost(self, state, action):
        return Pendulum.angle_normalize(state[:, 0]) ** 2 + self.w_thdot * state[:, 1] ** 2


class CudaModel(AnalyticModel):
    """A user-written analytic model compiled INTO the fused kernel.

    The reference takes arbitrary Python callables (mppi.py:63-64); those cannot be inlined into CUDA, so for
    analytic dynamics/costs you give the two function bodies as CUDA C++ and the engine compiles the fused / split-cost /
    resident / states kernels for your model at run time with NVRTC (in process, cached on disk: pytorch_mppi_b200.rtc)
    and loads them into the stock library (`mppi_user_model_register`).  Inside the bodies:

        x[NX]      state (read/write in `step_code`, read-only in `cost_code` / `terminal_code`)
        u[NU]      action, already multiplied by u_scale
        p[...]     your `params` (cast to the controller dtype)
        real       the controller dtype; O::add/sub/mul/div are individually rounded ops (use them if you want
                   bit-level agreement with an eager fp32 torch implementation), sinf/exp/... via O::sin_, O::exp_

    `cost_code` and `terminal_code` must `return` a `real`.  `dynamics` / `running_cost` / `terminal_cost`
    are the torch callables of the same model (used by the stepped route and as your simulator;
    `get_rollouts` runs the compiled `step_code`); pass ``model.dynamics`` / ``model.running_cost`` to the
    controller as usual.
    """
    model_id = _cabi.MODEL_USER

    def __init__(self, nx, nu, step_code, cost_code, params=(), terminal_code=None, dynamics=None, running_cost=None,
                 terminal_cost=None):
        if not (1 <= nx <= _cabi.MPPI_MAX_NX and 1 <= nu <= _cabi.MPPI_MAX_NU):
            raise ValueError(f"nx <= {_cabi.MPPI_MAX_NX} and nu <= {_cabi.MPPI_MAX_NU} required")
        self.nx, self.nu = int(nx), int(nu)
        self.params = [float(v) for v in params]
        self._step_code, self._cost_code, self._terminal_code = step_code, cost_code, terminal_code
        self._dyn, self._cost, self._term = dynamics, running_cost, terminal_cost
        self._lib_path = None
        self._rtc_handles = {}

    @property
    def has_terminal(self):
        return self._terminal_code is not None

    def param_blob(self):
        return self.params[: _cabi.MPPI_MODEL_PARAM_DOUBLES]

    def param_blob_ext(self):
        return self.params[_cabi.MPPI_MODEL_PARAM_DOUBLES:]

    def header_text(self):
        n = max(len(self.params), 1)
        term = self._terminal_code if self._terminal_code is not None else "return (real)0;"
        return f"""// generated by pytorch_mppi_b200.models.CudaModel
namespace mppi {{
struct UserModel {{
    static const int NX = {self.nx}, NU = {self.nu}, NP = {n};
    template <typename real> struct P {{ real v[NP]; }};
    template <typename real> static void load(P<real>& P_, const double* b, const double* ext = nullptr, int n_ext = 0) {{
        for (int i = 0; i < NP; ++i)
            P_.v[i] = (real)(i < {_cabi.MPPI_MODEL_PARAM_DOUBLES} ? b[i] : (ext != nullptr && i - {_cabi.MPPI_MODEL_PARAM_DOUBLES} < n_ext ? ext[i - {_cabi.MPPI_MODEL_PARAM_DOUBLES}] : 0.0));
    }}
    template <typename real> static MPPI_HD void step(const P<real>& P_, real* x, const real* u) {{
        typedef Ops<real> O;
        const real* p = P_.v;
        (void)p;
        {self._step_code}
    }}
    template <typename real> static MPPI_HD real cost(const P<real>& P_, const real* x, const real* u) {{
        typedef Ops<real> O;
        const real* p = P_.v;
        (void)p; (void)u;
        {self._cost_code}
    }}
    template <typename real> static MPPI_HD bool has_terminal(const P<real>&) {{ return {"true" if self.has_terminal else "false"}; }}
    template <typename real> static MPPI_HD real terminal(const P<real>& P_, const real* x) {{
        typedef Ops<real> O;
        const real* p = P_.v;
        (void)p; (void)x;
        {term}
    }}
}};
}}  // namespace mppi
"""

    def compile_rtc(self, dtype, variant):
        """(cubin bytes, lowered kernel names) of this model's kernels for one dtype and controller variant
        (0 MPPI / 1 SMPPI / 2 KMPPI): NVRTC, cached on disk.  Needs no GPU."""
        from . import rtc
        return rtc.compile_user_model(self.header_text(), "float" if dtype == torch.float32 else "double", int(variant))

    def rtc_handle(self, lib, dtype, variant):
        """The handle `MppiFusedParams.user_model` takes: the compiled kernels registered with the C library on the
        CURRENT CUDA device (once per (library, dtype, variant, device))."""
        import ctypes as C
        key = (id(lib), dtype, int(variant), torch.cuda.current_device())
        if key not in self._rtc_handles:
            cubin, names = self.compile_rtc(dtype, variant)
            arr = (C.c_char_p * len(names))(*[None if n is None else n.encode() for n in names])
            handle = C.c_void_p()
            buf = C.create_string_buffer(cubin, len(cubin))
            rc = lib.mppi_user_model_register(buf, len(cubin), self.nx, self.nu, max(len(self.params), 1),
                                              _cabi.F32 if dtype == torch.float32 else _cabi.F64, int(variant), arr, C.byref(handle))
            _cabi.check(rc, "mppi_user_model_register")
            self._rtc_handles[key] = handle
        return self._rtc_handles[key]

    def library_path(self):
        """The nvcc route (MPPI_B200_USER_MODEL_BUILD=nvcc): path of a (cached) variant library with this model linked in;
        builds it on first use.  Needs the CUDA toolkit."""
        if self._lib_path is None:
            from . import build
            self._lib_path = build.build_user_model(self.header_text())
        return self._lib_path

    def dynamics(self, state, action):
        if self._dyn is None:
            raise NotImplementedError("this CudaModel was created without a torch `dynamics` callable")
        return self._dyn(state, action)

    def running_cost(self, state, action):
        if self._cost is None:
            raise NotImplementedError("this CudaModel was created without a torch `running_cost` callable")
        return self._cost(state, action)

    def terminal_cost(self, states, actions):
        if self._term is None:
            raise NotImplementedError("this CudaModel was created without a torch `terminal_cost` callable")
        return self._term(states, actions)


def resolve_fused_model(dynamics, running_cost, terminal_state_cost) -> Optional[AnalyticModel]:
    """The registered model these plugins belong to, or None if they are not (all) bound methods of
    one registered model — in which case the controller uses the per-step path."""
    m = AnalyticModel.owner_of(dynamics)
    if m is None or AnalyticModel.owner_of(running_cost) is not m:
        return None
    # the class that REGISTERED the model (declares model_id): a user subclass that overrides dynamics / running_cost /
    # terminal_cost in Python is no longer the compiled model and must take the stepped route
    reg = next((c for c in type(m).__mro__ if "model_id" in c.__dict__ and c is not AnalyticModel), None)
    if reg is None:
        return None
    if getattr(dynamics, "__func__", None) is not reg.__dict__.get("dynamics"):
        return None
    if getattr(running_cost, "__func__", None) is not reg.__dict__.get("running_cost"):
        return None
    if terminal_state_cost is None:
        # the kernel adds the model's terminal cost iff terminal_scale != 0; without the plugin the
        # reference adds none, so only models whose terminal term is off qualify
        return m if not m.has_terminal else None
    if AnalyticModel.owner_of(terminal_state_cost) is not m or not m.has_terminal:
        return None
    if getattr(terminal_state_cost, "__func__", None) is not reg.__dict__.get("terminal_cost"):
        return None
    return m
