"""ctypes binding of the CPU oracle (oracle/blang_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(blangsdk_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAM_NORMAL, FAM_LOGISTIC, FAM_MIXTURE, FAM_ISING, FAM_BETABINOMIAL = 1, 2, 3, 4, 5
INIT_COPIES, INIT_FORWARD = 0, 1
ORDER_REFERENCE, ORDER_CHECKERBOARD = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class PtConfig(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("use_prior_samples", C.c_int32), ("reversible", C.c_int32),
                ("anneal_support", C.c_int32), ("n_threads", C.c_int32), ("sweep_order", C.c_int32),
                ("n_passes_per_scan", C.c_double), ("seed", C.c_uint64), ("replica", C.c_uint32)]


class RoundStats(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("n_scans_in_round", C.c_int32),
                ("swap_accept_mean", _dp), ("energy_mean", _dp), ("energy_var", _dp), ("energy_cov", _dp),
                ("log_sum", _dp), ("log_sum_n", _lp), ("round_trips", C.c_int64), ("Lambda", C.c_double),
                ("log_z_stepping_stone", C.c_double), ("log_z_thermo", C.c_double)]


class RunSummary(C.Structure):
    _fields_ = [("n_rounds", C.c_int32), ("round_n_scans", C.c_int32 * 64), ("round_is_adapt", C.c_int32 * 64),
                ("round_Lambda", C.c_double * 64), ("round_log_z", C.c_double * 64),
                ("round_trips", C.c_int64 * 64), ("round_min_accept", C.c_double * 64),
                ("round_mean_accept", C.c_double * 64)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "blang_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    vp = C.c_void_p
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    for f in ("orc_lngamma", "orc_lngamma_pike_hill", "orc_logistic"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double]
    L.orc_set_lngamma.argtypes = [C.c_int]
    L.orc_log_add.restype = C.c_double
    L.orc_log_add.argtypes = [C.c_double, C.c_double]
    L.orc_model_normal.restype = vp
    L.orc_model_normal.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_model_logistic.restype = vp
    L.orc_model_logistic.argtypes = [_dp, _ip, C.c_int, C.c_int, C.c_double]
    L.orc_model_mixture.restype = vp
    L.orc_model_mixture.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 5
    L.orc_model_ising.restype = vp
    L.orc_model_ising.argtypes = [C.c_int, C.c_double, C.c_double]
    L.orc_model_mixture_indicators.restype = vp
    L.orc_model_mixture_indicators.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 5
    L.orc_model_changepoint.restype = vp
    L.orc_model_changepoint.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_model_anneal_pair.restype = vp
    L.orc_model_anneal_pair.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.orc_model_betabinomial.restype = vp
    L.orc_model_betabinomial.argtypes = [_ip, _ip, C.c_int, C.c_double]
    L.orc_pt_error.restype = C.c_int
    L.orc_pt_error.argtypes = [vp]
    L.orc_model_rockets_latent.restype = vp
    L.orc_model_rockets_latent.argtypes = [_ip, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_model_linreg.restype = vp
    L.orc_model_linreg.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double]
    L.orc_model_hierarchical.restype = vp
    L.orc_model_hierarchical.argtypes = [_ip, _ip, C.c_int, C.c_double, C.c_double]
    L.orc_model_iid.restype = vp
    L.orc_model_iid.argtypes = [C.c_int, _dp, _dp, C.c_int, _ip, _dp, _dp]
    L.orc_model_free.argtypes = [vp]
    for f in ("orc_model_n_reals", "orc_model_n_ints", "orc_model_n_samplers", "orc_model_n_factors"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_int
    L.orc_model_log_density.argtypes = [vp, _dp, _ip, C.c_double, C.c_int, _dp]
    L.orc_model_factor_value.restype = C.c_double
    L.orc_model_factor_value.argtypes = [vp, C.c_int, _dp, _ip]
    L.orc_model_factor_is_annealed.argtypes = [vp, C.c_int]
    L.orc_pt_create.restype = vp
    L.orc_pt_create.argtypes = [vp, C.POINTER(PtConfig)]
    L.orc_pt_free.argtypes = [vp]
    L.orc_ladder_equally_spaced.argtypes = [C.c_int, _dp]
    L.orc_pt_set_ladder.argtypes = [vp, _dp, C.c_int]
    L.orc_pt_get_ladder.argtypes = [vp, _dp]
    L.orc_pt_set_state.argtypes = [vp, C.c_int, _dp, _ip]
    L.orc_pt_get_state.argtypes = [vp, C.c_int, _dp, _ip]
    L.orc_pt_initialize.argtypes = [vp, C.c_int]
    L.orc_pt_initialize_scm.restype = C.c_int
    L.orc_pt_initialize_scm.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp, _ip, _ip]
    L.orc_pt_densities.argtypes = [vp, _dp, _dp, _ip, _dp]
    L.orc_pt_check_caches.restype = C.c_double
    L.orc_pt_check_caches.argtypes = [vp]
    L.orc_pt_scan.argtypes = [vp, _dp, _ip]
    L.orc_pt_move_kernel.argtypes = [vp]
    L.orc_pt_swap_kernel.argtypes = [vp, _ip]
    for f in ("orc_pt_n_evals", "orc_pt_n_factor_evals", "orc_pt_n_scans"):
        getattr(L, f).restype = C.c_uint64
        getattr(L, f).argtypes = [vp]
    L.orc_pt_round_stats.argtypes = [vp, C.POINTER(RoundStats)]
    L.orc_pt_adapt.argtypes = [vp, _dp]
    L.orc_rounds.restype = C.c_int
    L.orc_rounds.argtypes = [C.c_int, C.c_double, _ip, _ip]
    L.orc_pt_perform_inference.restype = C.c_int
    L.orc_pt_perform_inference.argtypes = [vp, C.c_int, C.c_double, _dp, _ip, _dp, _ip, C.POINTER(RunSummary)]
    L.orc_monotone_spline_build.argtypes = [_dp, _dp, C.c_int, _dp]
    L.orc_monotone_spline_eval.restype = C.c_double
    L.orc_monotone_spline_eval.argtypes = [_dp, _dp, _dp, C.c_int, C.c_double]
    L.orc_adapt_ladder.argtypes = [_dp, _dp, C.c_int, _dp, _dp]
    L.orc_ising_exact_log_z.restype = C.c_double
    L.orc_ising_exact_log_z.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    _LIB = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def rounds(n_scans, adapt_fraction):
    rn = (C.c_int32 * 64)()
    ra = (C.c_int32 * 64)()
    k = lib().orc_rounds(n_scans, adapt_fraction, rn, ra)
    return [(int(rn[i]), bool(ra[i])) for i in range(k)]


def ladder_equally_spaced(n):
    out = np.zeros(n)
    lib().orc_ladder_equally_spaced(n, _d(out))
    return out


def adapt_ladder(betas_asc, accept_asc):
    b = np.ascontiguousarray(betas_asc, dtype=np.float64)
    a = np.ascontiguousarray(accept_asc, dtype=np.float64)
    out = np.zeros_like(b)
    cum = np.zeros_like(b)
    lib().orc_adapt_ladder(_d(b), _d(a), len(b), _d(out), _d(cum))
    return out, cum


IID_DISTS = dict(poisson=1, binomial=2, negbinomial=3, studentt=4, geometric=5, laplace=6, exponential=7, gamma=8,
                 weibull=9, gumbel=10, logisticdist=11, halfstudentt=12, loglogistic=13, gompertz=14, yulesimon=15,
                 uniformmax=16)
IID_PRIORS = dict(normal=0, gamma=1, exponential=2, beta=3, uniform=4)


class Model:
    """One of the catalogue families, built from a blangsdk_b200.datasets spec dict."""

    def __init__(self, spec):
        L = lib()
        self.spec = spec
        fam = spec["family"]
        h = spec["hyper"]
        if fam == "normal":
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            self.h = L.orc_model_normal(_d(y), len(y), h["m0"], h["v0"], h["rate"])
        elif fam == "logistic":
            x = np.ascontiguousarray(spec["x"], dtype=np.float64)
            y = np.ascontiguousarray(spec["y"], dtype=np.int32)
            self.h = L.orc_model_logistic(_d(x), _i(y), x.shape[0], x.shape[1], h["v0"])
        elif fam == "mixture":
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            self.h = L.orc_model_mixture(_d(y), len(y), int(h["K"]), h["m0"], h["v0"], h["gshape"], h["grate"], h["conc"])
        elif fam == "mixture_indicators":
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            self.h = L.orc_model_mixture_indicators(_d(y), len(y), int(h["K"]), h["m0"], h["v0"], h["gshape"], h["grate"], h["conc"])
        elif fam == "changepoint":
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            self.h = L.orc_model_changepoint(_d(y), len(y), h["m0"], h["v0"], h["s2"])
        elif fam == "anneal_pair":
            self.h = L.orc_model_anneal_pair(float(spec["x"]), h["m0"], h["v0"], h["s2"], int(h["manual"]))
        elif fam == "ising":
            self.h = L.orc_model_ising(int(h["N"]), h["coupling"], h["moment"])
        elif fam == "betabinomial":
            y = np.ascontiguousarray(spec["y"], dtype=np.int32)
            nt = np.ascontiguousarray(spec["ntrials"], dtype=np.int32)
            self.h = L.orc_model_betabinomial(_i(y), _i(nt), len(y), h["rate"])
        elif fam == "rockets_latent":
            y = np.ascontiguousarray(spec["y"], dtype=np.int32)
            self.h = L.orc_model_rockets_latent(_i(y), len(y), h["rate_a"], h["rate_b"], h["pois_mean"])
        elif fam == "linreg":
            x = np.ascontiguousarray(spec["x"], dtype=np.float64)
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            self.h = L.orc_model_linreg(_d(x), _d(y), x.shape[0], x.shape[1], h["v0"], h["smax"])
        elif fam == "hierarchical":
            y = np.ascontiguousarray(spec["y"], dtype=np.int32)
            nt = np.ascontiguousarray(spec["ntrials"], dtype=np.int32)
            self.h = L.orc_model_hierarchical(_i(y), _i(nt), len(y), h["rate_a"], h["rate_b"])
        elif fam == "iid":
            y = np.ascontiguousarray(spec["y"], dtype=np.float64)
            tr = np.ascontiguousarray(spec["trials"], dtype=np.float64) if spec.get("trials") is not None else None
            kinds = np.array([IID_PRIORS[p[0]] for p in h["priors"]], dtype=np.int32)
            pa = np.array([float(p[1]) for p in h["priors"]], dtype=np.float64)
            pb = np.array([float(p[2]) if len(p) > 2 else 0.0 for p in h["priors"]], dtype=np.float64)
            self.h = L.orc_model_iid(IID_DISTS[h["dist"]], _d(y), _d(tr), len(y), _i(kinds), _d(pa), _d(pb))
            if not self.h:
                raise ValueError("iid: unknown distribution")
        else:
            raise ValueError(fam)
        self.n_reals = L.orc_model_n_reals(self.h)
        self.n_ints = L.orc_model_n_ints(self.h)
        self.n_samplers = L.orc_model_n_samplers(self.h)
        self.n_factors = L.orc_model_n_factors(self.h)

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def log_density(self, reals=None, ints=None, beta=1.0, anneal_support=True):
        """-> dict(logdens, loglik, noos, fixed) from a from-scratch sum over every factor."""
        re = None if reals is None else np.ascontiguousarray(reals, dtype=np.float64)
        it = None if ints is None else np.ascontiguousarray(ints, dtype=np.int32)
        out = np.zeros(4)
        lib().orc_model_log_density(self.h, _d(re), _i(it), float(beta), int(anneal_support), _d(out))
        return dict(logdens=out[0], loglik=out[1], noos=int(out[2]), fixed=out[3])

    def factor_value(self, f, reals=None, ints=None):
        re = None if reals is None else np.ascontiguousarray(reals, dtype=np.float64)
        it = None if ints is None else np.ascontiguousarray(ints, dtype=np.int32)
        return lib().orc_model_factor_value(self.h, f, _d(re), _i(it))

    def factor_is_annealed(self, f):
        return bool(lib().orc_model_factor_is_annealed(self.h, f))


class PT:
    """The reference PT engine restated (ParallelTempering.java + factories/PT.java)."""

    def __init__(self, model, nChains=8, seed=1, nPassesPerScan=3.0, usePriorSamples=True, reversible=False,
                 annealSupport=True, nThreads=1, sweepOrder=ORDER_REFERENCE, replica=0):
        self.model = model
        self.n = nChains
        cfg = PtConfig(nChains, int(usePriorSamples), int(reversible), int(annealSupport), int(nThreads),
                       int(sweepOrder), float(nPassesPerScan), int(seed), int(replica))
        self.h = lib().orc_pt_create(model.h, C.byref(cfg))

    def __del__(self):
        try:
            lib().orc_pt_free(self.h)
        except Exception:
            pass

    def set_ladder(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64)
        lib().orc_pt_set_ladder(self.h, _d(b), len(b))

    def get_ladder(self):
        out = np.zeros(self.n)
        lib().orc_pt_get_ladder(self.h, _d(out))
        return out

    def set_state(self, pos, reals=None, ints=None):
        re = None if reals is None else np.ascontiguousarray(reals, dtype=np.float64)
        it = None if ints is None else np.ascontiguousarray(ints, dtype=np.int32)
        lib().orc_pt_set_state(self.h, pos, _d(re), _i(it))

    def get_state(self, pos):
        re = np.zeros(max(self.model.n_reals, 1))
        it = np.zeros(max(self.model.n_ints, 1), dtype=np.int32)
        lib().orc_pt_get_state(self.h, pos, _d(re), _i(it))
        return re[:self.model.n_reals], it[:self.model.n_ints]

    def initialize(self, init_type=INIT_FORWARD):
        lib().orc_pt_initialize(self.h, init_type)

    def initialize_scm(self, nParticles=100, threshold=0.9, resamplingESSThreshold=0.5):
        ln, nt, nr = C.c_double(), C.c_int32(), C.c_int32()
        rc = lib().orc_pt_initialize_scm(self.h, nParticles, threshold, resamplingESSThreshold, C.byref(ln), C.byref(nt), C.byref(nr))
        if rc != 0:
            raise RuntimeError("Invalid logDensity ratio (0/0)")
        return dict(logNormEstimate=ln.value, nTemperatures=nt.value, nResamplings=nr.value)

    def densities(self):
        ld, ll, fx = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n)
        no = np.zeros(self.n, dtype=np.int32)
        lib().orc_pt_densities(self.h, _d(ld), _d(ll), _i(no), _d(fx))
        return dict(logdens=ld, loglik=ll, noos=no, fixed=fx)

    def check_caches(self):
        return lib().orc_pt_check_caches(self.h)

    def scan(self):
        e = np.zeros(self.n)
        ind = np.zeros(self.n, dtype=np.int32)
        lib().orc_pt_scan(self.h, _d(e), _i(ind))
        return e, ind

    def error(self):
        """sticky bits of conditions that throw in the Java engine (ORC_ERR_* in blang_oracle.h)"""
        return lib().orc_pt_error(self.h)

    def n_evals(self):
        return int(lib().orc_pt_n_evals(self.h))

    def n_factor_evals(self):
        return int(lib().orc_pt_n_factor_evals(self.h))

    def round_stats(self):
        n = self.n
        arrs = dict(swap_accept_mean=np.zeros(n), energy_mean=np.zeros(n), energy_var=np.zeros(n),
                    energy_cov=np.zeros(n), log_sum=np.zeros(n))
        lsn = np.zeros(n, dtype=np.int64)
        st = RoundStats()
        for k, v in arrs.items():
            setattr(st, k, _d(v))
        st.log_sum_n = lsn.ctypes.data_as(_lp)
        lib().orc_pt_round_stats(self.h, C.byref(st))
        out = {k: (v[:n - 1] if k in ("swap_accept_mean", "log_sum") else v) for k, v in arrs.items()}
        out.update(log_sum_n=lsn[:n - 1], round_trips=int(st.round_trips), Lambda=st.Lambda,
                   log_z_stepping_stone=st.log_z_stepping_stone, log_z_thermo=st.log_z_thermo,
                   n_scans_in_round=int(st.n_scans_in_round))
        return out

    def adapt(self):
        knots = np.zeros(2 * self.n)
        lib().orc_pt_adapt(self.h, _d(knots))
        return knots[:self.n], knots[self.n:]

    def perform_inference(self, nScans=1000, adaptFraction=0.5, want_samples=True):
        total = sum(r[0] for r in rounds(nScans, adaptFraction if self.n > 1 else 0.0))
        nr, ni = self.model.n_reals, self.model.n_ints
        samples = np.zeros((total, nr)) if (want_samples and nr) else None
        isamples = np.zeros((total, ni), dtype=np.int32) if (want_samples and ni) else None
        energies = np.zeros((total, self.n))
        ind = np.zeros((total, self.n), dtype=np.int32)
        summ = RunSummary()
        lib().orc_pt_perform_inference(self.h, nScans, adaptFraction, _d(samples), _i(isamples), _d(energies),
                                       _i(ind), C.byref(summ))
        k = summ.n_rounds
        rounds_out = [dict(nScans=int(summ.round_n_scans[i]), isAdapt=bool(summ.round_is_adapt[i]),
                           Lambda=summ.round_Lambda[i], logZ=summ.round_log_z[i],
                           roundTrips=int(summ.round_trips[i]), minAccept=summ.round_min_accept[i],
                           meanAccept=summ.round_mean_accept[i]) for i in range(k)]
        return dict(samples=samples, int_samples=isamples, energies=energies, indicators=ind, rounds=rounds_out)
