"""Built-in device integrands.

A Julia closure `f(x, c)` (main.jl:25-33) cannot run on the GPU without CUDA.jl codegen, which the
north star excludes, so an integrand is a device function compiled into libmcib200.so and selected
by id.  Each entry fixes the number of integrands it returns and the variable layout it reads.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

from . import _abi


@dataclass(frozen=True)
class Integrand:
    name: str
    id: int
    n_integrand: int
    userdata: tuple = ()
    doc: str = ""
    source: str | None = None  # CUDA C++ source of a user integrand (id == F_USER)
    complex: bool = False      # ComplexF64 values: the device function writes (re, im) pairs

    def with_userdata(self, data: Sequence[float]) -> "Integrand":
        return Integrand(self.name, self.id, self.n_integrand, tuple(float(x) for x in data), self.doc, self.source,
                         self.complex)


def user_integrand(source: str, n_integrand: int = 1, userdata: Sequence[float] = (), name: str = "user",
                   complex: bool = False) -> Integrand:
    """The reference takes any closure f(x, c) (main.jl:25-33); here the integrand is CUDA C++ source compiled by NVRTC
    at engine creation into a kernel specialised to the problem.  `source` must define

        template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w);

    X(leaf, pos): value of pool slot `pos` (1-based, the reference's x[pos]) of leaf variable `leaf` (0-based, top-level
    order then leaf order); w[0..n_integrand-1] receives the values.  complex=True (with Configuration(type=complex)):
    w[2i], w[2i+1] = real and imaginary part of integrand i."""
    return Integrand(name, _abi.F_USER, int(n_integrand), tuple(float(x) for x in userdata), "user source (NVRTC)", source,
                     bool(complex))


@dataclass(frozen=True)
class UserMeasure:
    """The `measure` kwarg of integrate (main.jl:35-37) as CUDA C++ source defining

        template <class XA> __device__ void mci_user_measure(const XA &X, double *obs, const double *relw, const double *userdata);

    obs = the flattened config.observable (integrand i at the running offset of the observable lengths);
    relw[i] = relative weight of integrand i (under :mcmc zero except for the current integrand)."""
    source: str


def user_measure(source: str) -> UserMeasure:
    return UserMeasure(source)


gauss_nd = Integrand("gauss_nd", _abi.F_GAUSS_ND, 1, doc="(2π)^(-d/2) exp(-|x|²/2); one Continuous pool, dof=[[d]]")
log_over_sqrt = Integrand("log_over_sqrt", _abi.F_LOG_OVER_SQRT, 1, doc="log(x)/√x (test/montecarlo.jl:129-134)")
sphere1 = Integrand("sphere1", _abi.F_SPHERE1, 1, doc="x₁²+x₂²<1 (test/montecarlo.jl:4-9)")
sphere2 = Integrand("sphere2", _abi.F_SPHERE2, 2, doc="circle and sphere indicators, dof=[[2],[3]] (test/montecarlo.jl:19-49)")
discrete_x = Integrand("discrete_x", _abi.F_DISCRETE_X, 1, doc="f = x[1] on a Discrete pool (test/montecarlo.jl:104-111)")
discrete2_const = Integrand("discrete2_const", _abi.F_DISCRETE2_CONST, 1, doc="f = 1 on Discrete([(1,3),(1,4)]) (test/montecarlo.jl:113-120)")
lattice_pool = Integrand("lattice_pool", _abi.F_LATTICE_POOL, 1, doc="1/(1-cos x cos y cos z)/π³, pool dof=[[3]] (test/montecarlo.jl:136-147)")
lattice_comp = Integrand("lattice_comp", _abi.F_LATTICE_COMP, 1, doc="same on CompositeVar(X,Y,Z), dof=1 (test/montecarlo.jl:149-180)")
hypersphere3 = Integrand("hypersphere3", _abi.F_HYPERSPHERE3, 3, doc="nested unit balls d=2,3,4 (test/montecarlo.jl:200-216, N=3)")
poly3 = Integrand("poly3", _abi.F_POLY3, 3, doc="(x, x², x³) (test/thread.jl)")
x2_plus_y2 = Integrand("x2_plus_y2", _abi.F_X2_PLUS_Y2, 1, doc="x₁²+x₂² (main.jl:64)")
complex1 = Integrand("complex1", _abi.F_COMPLEX1, 1, doc="x + x² im (test/montecarlo.jl:166-170, TestComplex1)", complex=True)
complex2 = Integrand("complex2", _abi.F_COMPLEX2, 2, doc="(x, x² im), dof=[[1],[1]] (test/montecarlo.jl:172-198, TestComplex2)",
                     complex=True)


def genz_corner_peak(d: int, layout: str = "pool", csum: float = 1.85) -> Integrand:
    """(1+Σ cᵢxᵢ)^-(d+1) with cᵢ = csum·i/Σi (d=16 → cᵢ = 1.85·i/136)."""
    tot = d * (d + 1) // 2
    c = tuple(csum * i / tot for i in range(1, d + 1))
    fid = _abi.F_GENZ_CORNER_POOL if layout == "pool" else _abi.F_GENZ_CORNER_COMP
    return Integrand(f"genz_corner_{layout}{d}", fid, 1, c)


def genz_corner_peak_exact(c: Sequence[float]) -> float:
    """Closed form by inclusion–exclusion: ∫(1+c·x)^-(d+1) = Σ_S (-1)^|S| / (1+Σ_{i∈S}c_i) / (d! Πc)."""
    from fractions import Fraction
    import itertools

    d = len(c)
    cf = [Fraction(x).limit_denominator(10**15) for x in c]
    tot = Fraction(0)
    for r in range(d + 1):
        for S in itertools.combinations(range(d), r):
            tot += Fraction((-1) ** r) / (1 + sum(cf[i] for i in S))
    prod = Fraction(1)
    for x in cf:
        prod *= x
    return float(tot / (math.factorial(d) * prod))


def product4(a=(0.25, 0.5, 0.75, 1.0)) -> Integrand:
    """f_k = Π_i (1 + a_k (x_i − ½)), four integrands, each integrates to 1."""
    return Integrand("product4", _abi.F_PRODUCT_4, 4, tuple(float(x) for x in a))


@dataclass
class BubblePara:
    """example/bubble.jl:10-22 (Base.@kwdef struct Para)."""
    rs: float = 1.0
    beta: float = 25.0
    spin: int = 2
    Qsize: int = 4
    dim: int = 3
    me: float = 0.5
    kF: float = field(init=False)
    extQ: list = field(init=False)
    β: float = field(init=False)

    def __post_init__(self):
        self.kF = (9 * math.pi / (2 * self.spin)) ** (1 / 3) / self.rs if self.dim == 3 else math.sqrt(4 / self.spin) / self.rs
        n = self.Qsize
        # LinRange(0.0*kF, 1.5*kF, Qsize)
        self.extQ = [((1 - i / (n - 1)) * 0.0 * self.kF + (i / (n - 1)) * 1.5 * self.kF) for i in range(n)]
        self.β = self.beta / (self.kF ** 2 / 2 / self.me)

    def lindhard(self, q: float) -> float:
        """example/bubble.jl:24-38 (T=0 free-electron polarisation)."""
        me, kF, spin = self.me, self.kF, self.spin
        density = me * kF / (2 * math.pi ** 2)
        if q < 1e-6:
            q = 1e-6
        x = q / 2 / kF
        if abs(q - 2 * kF) > 1e-6:
            Pi = 1 + (1 - x ** 2) * math.log1p(4 * x / ((1 - x) ** 2)) / 4 / x
        else:
            Pi = 1.0
        return -Pi * density * spin / 2


def lindhard_finite_temperature(para: "BubblePara", q: float) -> float:
    """Static polarisation of example/bubble.jl at ITS temperature: the Maldague average
    Pi_T(q; mu) = int_0^inf dmu' Pi_0(q; mu') / (4T cosh^2((mu - mu') / 2T)) of the T=0 Lindhard function
    (example/bubble.jl:24-38) with mu = kF^2/2m fixed, as the integrand's dispersion has it (bubble.jl:62,68).
    The reference compares with the T=0 value under loose sigma ratios; this is the exact target (needs scipy)."""
    from scipy.integrate import quad

    me, kF, spin = para.me, para.kF, para.spin
    EF, T = kF ** 2 / (2 * me), 1.0 / para.β

    def pi0(kf):
        if kf <= 0:
            return 0.0
        qq = max(q, 1e-6)
        x = qq / 2 / kf
        Pi = 1 + (1 - x ** 2) * math.log(abs((1 + x) / (1 - x))) / 2 / x if abs(qq - 2 * kf) > 1e-9 else 1.0
        return -Pi * (me * kf / (2 * math.pi ** 2)) * spin / 2

    f = lambda mu: pi0(math.sqrt(2 * me * mu)) / (4 * T * math.cosh((mu - EF) / (2 * T)) ** 2)
    v, _ = quad(f, 0.0, EF + 40 * T, points=[EF, (max(q, 1e-6) / 2) ** 2 / (2 * me)], limit=400, epsabs=1e-13, epsrel=1e-12)
    return v


def bubble(para: BubblePara | None = None) -> Integrand:
    """example/bubble.jl:50-75; variables (R, θ, φ, T, Ext); userdata = [kF, β, me, spin, dim, Qsize, extQx…]."""
    para = para or BubblePara()
    ud = (para.kF, para.β, para.me, float(para.spin), float(para.dim), float(para.Qsize), *para.extQ)
    return Integrand("bubble", _abi.F_BUBBLE, 1, tuple(ud))


def bubble_fermik(para: BubblePara | None = None) -> Integrand:
    """test/bubble_FermiK.jl:54-75; variables (T, K::FermiK{3}, Ext), solver=:mcmc; userdata as `bubble`."""
    para = para or BubblePara()
    ud = (para.kF, para.β, para.me, float(para.spin), float(para.dim), float(para.Qsize), *para.extQ)
    return Integrand("bubble_fermik", _abi.F_BUBBLE_FERMIK, 1, tuple(ud))


CUBA11_SOURCE = r"""
// the 11 test integrands of the Cuba demo on the unit cube — the workload of the reference's own timing comparison,
// example/benchmark/cuba/benchmark.jl:35-46 (t1 .. t11) evaluated together like its test2 / test3 (:62-89)
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  const double x = X(0, 1), y = X(0, 2), z = X(0, 3);
  const double rsq = x * x + y * y + z * z, xyz = x * y * z;
  w[0] = sin(x) * cos(y) * exp(z);
  w[1] = 1.0 / ((x + y) * (x + y) + 0.003) * cos(y) * exp(z);
  w[2] = 1.0 / (3.75 - cospi(x) - cospi(y) - cospi(z));
  w[3] = fabs(rsq - 0.125);
  w[4] = exp(-rsq);
  w[5] = 1.0 / (1.0 - xyz + 1e-10);
  w[6] = sqrt(fabs(x - y - z));
  w[7] = exp(-xyz);
  w[8] = x * x / (cos(x + y + z + 1.0) + 5.0);
  w[9] = (x > 0.5) ? 1.0 / sqrt(xyz + 1e-5) : sqrt(xyz);
  w[10] = (rsq < 1.0) ? 1.0 : 0.0;
}
"""
# what the reference prints for this workload (comments of example/benchmark/cuba/benchmark.jl:121-131,148-158; Cuba's own
# Vegas at 1e6 evaluations :104-114): (estimate, error)
CUBA11_REFERENCE = {
    "vegas": [(0.6662389036758806, 0.0009370223870430191), (5.256465932385138, 0.007319608002109763),
              (0.3075345398612458, 0.00024117145696156536), (0.8788561548719585, 0.001217419754053441),
              (0.4161189703927425, 0.00038591698396624104), (1.2031224696691054, 0.0014223459029237325),
              (0.7100206908822683, 0.0007923510194941002), (0.8912870247582633, 0.0007501319218471907),
              (0.08030537560622107, 0.00013380714402722626), (2.396395164639421, 0.0035658664551599913),
              (0.5228992089675735, 0.0008064389576126581)],
    "vegasmc": [(0.6636210504918344, 0.0020729178999968614), (5.242919616854019, 0.02260128007562855),
                (0.30681239353464046, 0.0007850307152916456), (0.8780127497538999, 0.0027696019511458045),
                (0.4144658711129951, 0.0011614492930177123), (1.2002718763309468, 0.003387144818775182),
                (0.7095634473688555, 0.002250824818469535), (0.8880588065576488, 0.0022868326414642133),
                (0.08033351807552817, 0.0002957920163406668), (2.417025425999495, 0.011011866855686713),
                (0.5200202669127822, 0.0016844693189224066)],
    "cuba_vegas_1e6": [(0.6646695631037952, 0.0005192264010635556), (5.268641443643625, 0.00937817147924112),
                       (0.30780794363958314, 0.00016915694878023873), (0.8773127802819499, 0.0007272020591230047),
                       (0.4165412720874263, 0.00026445560918383573), (1.2020166963125194, 0.0007653098425722113),
                       (0.7096235523946715, 0.000489330660608438), (0.8912187300170151, 0.00040985374702141484),
                       (0.08018533940010736, 8.303642721037487e-5), (2.3963590312091303, 0.0032631645439722893),
                       (0.5236360284653482, 0.0005842099358276612)],
    "wall_s": {"vegas": 0.246125, "vegasmc": 0.495360},  # @time of integrate(...; neval=1e5) — 10 iterations, one CPU thread
}


def cuba11() -> Integrand:
    """The 11-integrand Cuba demo set of example/benchmark/cuba/benchmark.jl — the one workload the reference itself
    publishes a wall time for — as a user-source integrand: var = Continuous(0, 1), dof = [[3]] * 11."""
    return user_integrand(CUBA11_SOURCE, n_integrand=11, name="cuba11")


REGISTRY = {
    f.name: f
    for f in (gauss_nd, log_over_sqrt, sphere1, sphere2, discrete_x, discrete2_const, lattice_pool, lattice_comp,
              hypersphere3, poly3, x2_plus_y2, complex1, complex2)
}


def resolve(f) -> Integrand:
    if isinstance(f, Integrand):
        return f
    if isinstance(f, str) and f in REGISTRY:
        return REGISTRY[f]
    if callable(f):
        raise TypeError(
            "mcintegration.jl_b200 runs the integrand on the GPU: pass a built-in integrand "
            "(mcintegration_jl_b200.integrands.*) instead of a Python callable; see INTEGRATION.md"
        )
    raise KeyError(f"unknown integrand {f!r}; known: {sorted(REGISTRY)}")
