"""ctypes binding of libblangpt.so + a Python mirror of the reference's PT engine options.

`PT` keeps the option names of blang.engines.internals.factories.PT
(R/engines/internals/factories/PT.java:47-82) and blang.engines.ParallelTempering
(R/engines/ParallelTempering.java:25-40) so a test reads like the reference's own:

    pt = PT(nChains=8, nScans=1000, adaptFraction=0.5, random=1)          # initialization="SCM" like the reference
    pt.setSampledModel(datasets.normal_meanvar())
    result = pt.performInference()

Everything numeric happens inside the CUDA library; this module only marshals
pointers.  There is no CPU fallback: if the library or a GPU is missing the
constructor raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLANGPT_LIB") or os.path.join(_HERE, "libblangpt.so")   # BLANGPT_LIB: A/B of kernel builds (tools/build_variant.sh)

FAMILIES = {"normal": 1, "logistic": 2, "mixture": 3, "ising": 4, "betabinomial": 5, "iid": 6, "hierarchical": 7, "linreg": 8, "rockets_latent": 9,
            "mixture_indicators": 10, "changepoint": 11, "anneal_pair": 12}
# family "iid" (include/blangpt.h BLANGPT_IID_* / BLANGPT_PRIOR_*): observation laws and priors
IID_DISTS = {"poisson": 1, "binomial": 2, "negbinomial": 3, "studentt": 4, "geometric": 5, "laplace": 6,
             "exponential": 7, "gamma": 8, "weibull": 9, "gumbel": 10, "logisticdist": 11, "halfstudentt": 12,
             "loglogistic": 13, "gompertz": 14, "yulesimon": 15, "uniformmax": 16}
IID_PRIORS = {"normal": 0, "gamma": 1, "exponential": 2, "beta": 3, "uniform": 4}
INIT = {"COPIES": 0, "FORWARD": 1, "SCM": 2}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class BlangPTError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference throws RuntimeException)."""


class Config(C.Structure):
    _fields_ = [("nChains", C.c_int32), ("nReplicas", C.c_int32), ("random", C.c_uint64),
                ("nPassesPerScan", C.c_double), ("usePriorSamples", C.c_int32), ("reversible", C.c_int32),
                ("annealSupport", C.c_int32), ("treatNaNAsNegativeInfinity", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("worldSize", C.c_int32), ("ctasPerChain", C.c_int32),
                ("replicaOffset", C.c_int64), ("statisticRecordedMaxChainIndex", C.c_int32),
                ("reserved", C.c_int32)]


class Array(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_elem", C.c_int64), ("dtype", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32)]


class ScanOutputs(C.Structure):
    _fields_ = [("energy", _dp), ("logDensity", _dp), ("nOutOfSupport", _ip), ("swapIndicators", _ip),
                ("samples", _dp), ("intSamples", _ip)]


class RoundStats(C.Structure):
    _fields_ = [("nScansInRound", C.c_int64), ("swapAcceptMean", _dp), ("energyMean", _dp),
                ("energyVariance", _dp), ("energyCovariance", _dp), ("logSum", _dp), ("logSumN", _lp),
                ("roundTrips", _lp), ("Lambda", _dp), ("logZSteppingStone", _dp), ("logZThermodynamic", _dp)]


class Diagnostics(C.Structure):
    _fields_ = [("nScansInRound", C.c_int64), ("swapLowest", C.c_double), ("swapAverage", C.c_double),
                ("energyCorrelationAverage", C.c_double), ("energyCorrelationHighest", C.c_double),
                ("Lambda", C.c_double), ("inefficiency", C.c_double), ("timeToFirstRestart", C.c_double),
                ("effectiveNScans", C.c_double), ("temperedRestarts", C.c_int64), ("temperedRestartRate", C.c_double),
                ("asymptoticRoundTripCount", C.c_double), ("asymptoticRoundTripRate", C.c_double),
                ("nonAsymptoticRoundTripCount", C.c_double), ("nonAsymptoticRoundTripRate", C.c_double),
                ("logNormalizationSteppingStone", C.c_double), ("logNormalizationThermodynamic", C.c_double),
                ("cumulativeLambda", C.c_double * 99), ("lambdaInstantaneous", C.c_double * 99)]


class InferenceSummary(C.Structure):
    _fields_ = [("nRounds", C.c_int32), ("roundNScans", C.c_int32 * 64), ("roundIsAdapt", C.c_int32 * 64),
                ("Lambda", _dp), ("logZ", _dp), ("roundTrips", _lp), ("minAccept", _dp), ("meanAccept", _dp),
                ("roundMillis", C.c_double * 64)]


EXPORTS = [
    "blangpt_default_config", "blangpt_create", "blangpt_destroy", "blangpt_last_error", "blangpt_set_model",
    "blangpt_n_reals", "blangpt_n_ints", "blangpt_n_samplers", "blangpt_ctas_per_chain", "blangpt_kernel_variant", "blangpt_set_ladder", "blangpt_get_ladder",
    "blangpt_set_state", "blangpt_get_state", "blangpt_initialize", "blangpt_log_density", "blangpt_run_scans",
    "blangpt_get_round_stats", "blangpt_adapt", "blangpt_adapt_preview", "blangpt_reset_round", "blangpt_rounds",
    "blangpt_perform_inference", "blangpt_counters", "blangpt_set_stream", "blangpt_evaluate", "blangpt_prime",
    "blangpt_explore", "blangpt_table_ptr", "blangpt_local_chains", "blangpt_swap", "blangpt_synchronize",
    "blangpt_microbench", "blangpt_debug_logit_terms", "blangpt_make_ladder", "blangpt_get_diagnostics",
    "blangpt_set_scm_options", "blangpt_scm_summary", "blangpt_target_accept_ladder",
    "blangpt_peer_export", "blangpt_peer_connect", "blangpt_peer_mailbox", "blangpt_peer_connect_pointers",
]

_LIB = None


def load_library():
    """dlopen libblangpt.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise BlangPTError("libblangpt.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.blangpt_default_config.argtypes = [C.POINTER(Config)]
    L.blangpt_default_config.restype = None
    L.blangpt_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.blangpt_destroy.argtypes = [vp]
    L.blangpt_destroy.restype = None
    L.blangpt_last_error.argtypes = [vp]
    L.blangpt_last_error.restype = C.c_char_p
    L.blangpt_set_model.argtypes = [vp, C.c_int32, C.POINTER(Array), C.c_int32, _dp, C.c_int32]
    for f in ("blangpt_n_reals", "blangpt_n_ints", "blangpt_n_samplers", "blangpt_ctas_per_chain", "blangpt_kernel_variant"):
        getattr(L, f).argtypes = [vp]
    L.blangpt_set_ladder.argtypes = [vp, _dp, C.c_int64]
    L.blangpt_get_ladder.argtypes = [vp, _dp, C.c_int64]
    L.blangpt_set_state.argtypes = [vp, C.c_int64, C.c_int32, _dp, _ip]
    L.blangpt_get_state.argtypes = [vp, C.c_int64, C.c_int32, _dp, _ip]
    L.blangpt_initialize.argtypes = [vp, C.c_int32]
    L.blangpt_log_density.argtypes = [vp, _dp, _dp, _ip, _dp]
    L.blangpt_run_scans.argtypes = [vp, C.c_int32, C.POINTER(ScanOutputs)]
    L.blangpt_get_round_stats.argtypes = [vp, C.POINTER(RoundStats)]
    L.blangpt_adapt.argtypes = [vp, _dp]
    L.blangpt_adapt_preview.argtypes = [vp, _dp, _dp]
    L.blangpt_reset_round.argtypes = [vp]
    L.blangpt_rounds.argtypes = [C.c_int32, C.c_double, _ip, _ip]
    L.blangpt_perform_inference.argtypes = [vp, C.c_int32, C.c_double, C.c_int32, C.POINTER(ScanOutputs),
                                            C.POINTER(InferenceSummary)]
    L.blangpt_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.blangpt_set_stream.argtypes = [vp, vp]
    for f in ("blangpt_evaluate", "blangpt_prime", "blangpt_explore", "blangpt_synchronize"):
        getattr(L, f).argtypes = [vp]
    L.blangpt_table_ptr.argtypes = [vp, C.c_int32]
    L.blangpt_table_ptr.restype = vp
    L.blangpt_local_chains.argtypes = [vp, _lp]
    L.blangpt_local_chains.restype = C.c_int64
    L.blangpt_swap.argtypes = [vp, C.POINTER(ScanOutputs)]
    L.blangpt_microbench.argtypes = [C.c_int32, C.c_int32, _dp]
    L.blangpt_debug_logit_terms.argtypes = [C.c_int32, _dp, _ip, C.c_int64, _dp]
    L.blangpt_make_ladder.argtypes = [C.c_int32, C.c_int32, C.c_double, _dp, C.c_int32, C.c_int32, _dp]
    L.blangpt_get_diagnostics.argtypes = [vp, C.c_int64, C.POINTER(Diagnostics)]
    L.blangpt_set_scm_options.argtypes = [vp, C.c_int32, C.c_double, C.c_double]
    L.blangpt_scm_summary.argtypes = [vp, _dp, _ip, _ip]
    L.blangpt_target_accept_ladder.argtypes = [vp, C.c_int64, C.c_double, _dp, C.c_int32, _ip]
    L.blangpt_peer_export.argtypes = [vp, C.c_void_p]
    L.blangpt_peer_connect.argtypes = [vp, C.c_void_p, C.c_int32]
    L.blangpt_peer_mailbox.argtypes = [vp]
    L.blangpt_peer_mailbox.restype = vp
    L.blangpt_peer_connect_pointers.argtypes = [vp, C.POINTER(vp), _ip, C.c_int32]
    _LIB = L
    return L


def microbench(which, device=0):
    """Measured compute ceiling: which=0 fp64 FMA TFLOP/s, which=1 logistic row terms/s."""
    out = C.c_double()
    rc = load_library().blangpt_microbench(device, which, C.byref(out))
    if rc != 0:
        raise BlangPTError("microbench failed (no GPU?)")
    return out.value


LADDERS = {"EquallySpaced": 0, "Geometric": 1, "Polynomial": 2, "UserSpecified": 3}


def make_ladder(kind, nChains, param=None, annealingParameters=None, allowSplineGeneralization=False):
    """Initial ladders of R/engines/internals/ladders (EquallySpaced, Geometric(annealingScaling=0.8),
    Polynomial(power=3), UserSpecified(annealingParameters, allowSplineGeneralization)); descending."""
    if param is None:
        param = {"Geometric": 0.8, "Polynomial": 3.0}.get(kind, 0.0)
    user = None if annealingParameters is None else np.ascontiguousarray(annealingParameters, dtype=np.float64)
    out = np.zeros(max(1, nChains))
    rc = load_library().blangpt_make_ladder(LADDERS[kind], nChains, float(param), _d(user),
                                            0 if user is None else user.size, int(allowSplineGeneralization), _d(out))
    if rc != 0:
        raise BlangPTError("invalid ladder specification (%s, nChains=%d)" % (kind, nChains))
    return out[:nChains]


def ladder_from_another_exec(annealing_parameters_csv, nChains, allowSplineGeneralization=False):
    """ladders/FromAnotherExec.java:27-40: the final (isAdapt == false) round's schedule of an earlier run."""
    import csv
    from .results import open_table          # annealingParameters.csv or .csv.gz (FromAnotherExec.java:16)
    parsed = [0.0]
    with open_table(annealing_parameters_csv) as f:
        for row in csv.DictReader(f):
            if row["isAdapt"] == "false":
                parsed.append(float(row["value"]))
    return make_ladder("UserSpecified", nChains, annealingParameters=parsed,
                       allowSplineGeneralization=allowSplineGeneralization)


def debug_logit_terms(z, y, device=0):
    """Per-row Bernoulli(logistic(z)) terms as the CUDA kernels compute them (accuracy tests)."""
    z = np.ascontiguousarray(z, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.int32)
    out = np.zeros_like(z)
    rc = load_library().blangpt_debug_logit_terms(device, _d(z), _i(y), z.size, _d(out))
    if rc != 0:
        raise BlangPTError("debug_logit_terms failed")
    return out


def rounds(nScans, adaptFraction):
    """PT.rounds (PT.java:239-265) -> [(nScans, isAdapt), ...]"""
    rn = (C.c_int32 * 64)()
    ra = (C.c_int32 * 64)()
    k = load_library().blangpt_rounds(nScans, adaptFraction, rn, ra)
    if k < 0:
        raise BlangPTError("adaptFraction must be in [0,1)")
    return [(int(rn[i]), bool(ra[i])) for i in range(k)]


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class Engine:
    """Thin object wrapper over one blangpt_handle."""

    def __init__(self, nChains=8, nReplicas=1, random=1, nPassesPerScan=3.0, usePriorSamples=True,
                 reversible=False, annealSupport=True, treatNaNAsNegativeInfinity=False, device=0, rank=0,
                 worldSize=1, ctasPerChain=0, replicaOffset=0):
        self.L = load_library()
        cfg = Config()
        self.L.blangpt_default_config(C.byref(cfg))
        cfg.nChains, cfg.nReplicas, cfg.random = int(nChains), int(nReplicas), int(random)
        cfg.nPassesPerScan = float(nPassesPerScan)
        cfg.usePriorSamples, cfg.reversible = int(usePriorSamples), int(reversible)
        cfg.annealSupport, cfg.treatNaNAsNegativeInfinity = int(annealSupport), int(treatNaNAsNegativeInfinity)
        cfg.device, cfg.rank, cfg.worldSize = int(device), int(rank), int(worldSize)
        cfg.ctasPerChain, cfg.replicaOffset = int(ctasPerChain), int(replicaOffset)
        self.cfg = cfg
        self.N, self.R = int(nChains), int(nReplicas)
        self.h = C.c_void_p()
        rc = self.L.blangpt_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = C.c_void_p()
            raise BlangPTError(self.L.blangpt_last_error(None).decode())
        self.n_reals = self.n_ints = self.n_samplers = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.blangpt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BlangPTError(self.L.blangpt_last_error(self.h).decode())

    # ---- model ----
    def set_model(self, spec):
        fam = spec["family"]
        h = spec["hyper"]
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=np.float64 if dt == 0 else np.int32)
            keep.append(a)
            return Array(a.ctypes.data, a.size, dt, a.shape[0] if a.ndim else 1, a.shape[1] if a.ndim > 1 else 1)

        if fam == "normal":
            arrays, hyper = [arr(spec["y"], 0)], [h["m0"], h["v0"], h["rate"]]
        elif fam == "logistic":
            arrays, hyper = [arr(spec["x"], 0), arr(spec["y"], 1)], [h["v0"]]
        elif fam == "mixture":
            arrays = [arr(spec["y"], 0)]
            hyper = [float(h["K"]), h["m0"], h["v0"], h["gshape"], h["grate"], h["conc"]]
        elif fam == "mixture_indicators":
            arrays = [arr(spec["y"], 0)]
            hyper = [float(h["K"]), h["m0"], h["v0"], h["gshape"], h["grate"], h["conc"]]
        elif fam == "changepoint":
            arrays, hyper = [arr(spec["y"], 0)], [h["m0"], h["v0"], h["s2"]]
        elif fam == "anneal_pair":
            arrays, hyper = [], [float(spec["x"]), h["m0"], h["v0"], h["s2"], 1.0 if h["manual"] else 0.0]
        elif fam == "ising":
            arrays, hyper = [], [float(h["N"]), h["coupling"], h["moment"]]
        elif fam == "betabinomial":
            arrays, hyper = [arr(spec["y"], 1), arr(spec["ntrials"], 1)], [h["rate"]]
        elif fam == "rockets_latent":
            arrays, hyper = [arr(spec["y"], 1)], [h["rate_a"], h["rate_b"], h["pois_mean"]]
        elif fam == "linreg":
            arrays, hyper = [arr(spec["x"], 0), arr(spec["y"], 0)], [h["v0"], h["smax"]]
        elif fam == "hierarchical":
            arrays, hyper = [arr(spec["y"], 1), arr(spec["ntrials"], 1)], [h["rate_a"], h["rate_b"]]
        elif fam == "iid":
            if h["dist"] not in IID_DISTS:
                raise BlangPTError("iid: unsupported distribution %r" % (h["dist"],))
            arrays = [arr(spec["y"], 0)]
            if h["dist"] == "binomial":
                arrays.append(arr(spec["trials"], 0))
            hyper = [float(IID_DISTS[h["dist"]])]
            for p in h["priors"]:       # (kind, a[, b]) per parameter, declaration order
                if p[0] not in IID_PRIORS:
                    raise BlangPTError("iid: unsupported prior %r" % (p[0],))
                hyper += [float(IID_PRIORS[p[0]]), float(p[1]), float(p[2]) if len(p) > 2 else 0.0]
        else:
            raise BlangPTError("unsupported model family %r (closed catalogue; no CPU fallback)" % fam)
        ca = (Array * max(1, len(arrays)))(*arrays)
        hy = np.asarray(hyper, dtype=np.float64)
        self._check(self.L.blangpt_set_model(self.h, FAMILIES[fam], ca, len(arrays), _d(hy), len(hy)))
        self.n_reals = self.L.blangpt_n_reals(self.h)
        self.n_ints = self.L.blangpt_n_ints(self.h)
        self.n_samplers = self.L.blangpt_n_samplers(self.h)
        self.ctas_per_chain = self.L.blangpt_ctas_per_chain(self.h)
        self.kernel_variant = self.L.blangpt_kernel_variant(self.h)   # 0 = cluster per chain, 1 = pool
        self.family = fam

    # ---- multi-GPU exchange (worldSize > 1): connect the ranks' mailboxes; afterwards every method works as on one GPU ----
    def peer_handle(self):
        buf = C.create_string_buffer(64)
        self._check(self.L.blangpt_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, handles):
        """handles: the 64-byte strings of every rank's peer_handle(), in rank order"""
        blob = b"".join(handles)
        self._check(self.L.blangpt_peer_connect(self.h, C.c_char_p(blob), len(handles)))

    def set_ladder(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64).ravel()
        self._check(self.L.blangpt_set_ladder(self.h, _d(b), b.size))

    def get_ladder(self):
        out = np.zeros(self.R * self.N)
        self._check(self.L.blangpt_get_ladder(self.h, _d(out), out.size))
        return out.reshape(self.R, self.N)

    def set_state(self, position, reals=None, ints=None, replica=0):
        re = None if reals is None else np.ascontiguousarray(reals, dtype=np.float64)
        it = None if ints is None else np.ascontiguousarray(ints, dtype=np.int32)
        self._check(self.L.blangpt_set_state(self.h, replica, position, _d(re), _i(it)))

    def get_state(self, position, replica=0):
        re = np.zeros(max(1, self.n_reals))
        it = np.zeros(max(1, self.n_ints), dtype=np.int32)
        self._check(self.L.blangpt_get_state(self.h, replica, position, _d(re), _i(it)))
        return re[:self.n_reals], it[:self.n_ints]

    def initialize(self, initialization="FORWARD", nParticles=None, threshold=0.9, resamplingESSThreshold=0.5):
        if nParticles is not None:
            self._check(self.L.blangpt_set_scm_options(self.h, nParticles, threshold, resamplingESSThreshold))
        self._check(self.L.blangpt_initialize(self.h, INIT[initialization]))

    def scm_summary(self):
        ln, nt, nr = C.c_double(), C.c_int32(), C.c_int32()
        self._check(self.L.blangpt_scm_summary(self.h, C.byref(ln), C.byref(nt), C.byref(nr)))
        return dict(logNormEstimate=ln.value, nTemperatures=nt.value, nResamplings=nr.value)

    def log_density(self):
        M = self.R * self.N
        ld, ll, fx = np.zeros(M), np.zeros(M), np.zeros(M)
        no = np.zeros(M, dtype=np.int32)
        self._check(self.L.blangpt_log_density(self.h, _d(ld), _d(ll), _i(no), _d(fx)))
        sh = (self.R, self.N)
        return dict(logdens=ld.reshape(sh), loglik=ll.reshape(sh), noos=no.reshape(sh), fixed=fx.reshape(sh))

    def _outputs(self, n, want):
        R, N = self.R, self.N
        bufs = {}
        so = ScanOutputs()
        if "energy" in want:
            bufs["energy"] = np.zeros((n, R, N)); so.energy = _d(bufs["energy"])
        if "logDensity" in want:
            bufs["logDensity"] = np.zeros((n, R, N)); so.logDensity = _d(bufs["logDensity"])
        if "nOutOfSupport" in want:
            bufs["nOutOfSupport"] = np.zeros((n, R, N), dtype=np.int32); so.nOutOfSupport = _i(bufs["nOutOfSupport"])
        if "swapIndicators" in want:
            bufs["swapIndicators"] = np.zeros((n, R, N), dtype=np.int32); so.swapIndicators = _i(bufs["swapIndicators"])
        if "samples" in want and self.n_reals:
            bufs["samples"] = np.zeros((n, R, self.n_reals)); so.samples = _d(bufs["samples"])
        if "samples" in want and self.n_ints:
            bufs["intSamples"] = np.zeros((n, R, self.n_ints), dtype=np.int32); so.intSamples = _i(bufs["intSamples"])
        return so, bufs

    def run_scans(self, n, want=("energy", "swapIndicators", "samples")):
        so, bufs = self._outputs(n, want)
        self._check(self.L.blangpt_run_scans(self.h, n, C.byref(so) if want else None))
        return bufs

    def round_stats(self):
        R, N = self.R, self.N
        a = {k: np.zeros((R, N)) for k in ("swapAcceptMean", "energyMean", "energyVariance", "energyCovariance", "logSum")}
        lsn = np.zeros((R, N), dtype=np.int64)
        rt = np.zeros(R, dtype=np.int64)
        per = {k: np.zeros(R) for k in ("Lambda", "logZSteppingStone", "logZThermodynamic")}
        st = RoundStats()
        for k, v in a.items():
            setattr(st, k, _d(v))
        for k, v in per.items():
            setattr(st, k, _d(v))
        st.logSumN = lsn.ctypes.data_as(_lp)
        st.roundTrips = rt.ctypes.data_as(_lp)
        self._check(self.L.blangpt_get_round_stats(self.h, C.byref(st)))
        out = dict(a)
        out["swapAcceptMean"] = a["swapAcceptMean"][:, :N - 1]
        out["logSum"] = a["logSum"][:, :N - 1]
        out.update(per, logSumN=lsn[:, :N - 1], roundTrips=rt, nScansInRound=int(st.nScansInRound))
        return out

    def diagnostics(self, replica=0):
        """PT.reportParallelTemperingDiagnostics + reportLambdaFunctions for the current round."""
        d = Diagnostics()
        self._check(self.L.blangpt_get_diagnostics(self.h, replica, C.byref(d)))
        out = {k: getattr(d, k) for k, _ in Diagnostics._fields_ if k not in ("cumulativeLambda", "lambdaInstantaneous")}
        out["cumulativeLambda"] = np.array(d.cumulativeLambda)
        out["lambdaInstantaneous"] = np.array(d.lambdaInstantaneous)
        return out

    def target_accept_ladder(self, targetAccept, replica=0, capacity=65536):
        """resized ladder of PT.adapt's targetAccept branch (PT.java:156-163), descending"""
        out = np.zeros(capacity)
        n = C.c_int32()
        self._check(self.L.blangpt_target_accept_ladder(self.h, replica, float(targetAccept), _d(out), capacity, C.byref(n)))
        return out[:n.value].copy()

    def adapt(self):
        knots = np.zeros((self.R, 2, self.N))
        self._check(self.L.blangpt_adapt(self.h, _d(knots)))
        return knots

    def adapt_preview(self):
        """the ladder PT.adapt would install now (PT.java:103-107: the reference adapts once more after the last round
        "for instrumentation"), without touching the engine; returns (ladder[R][N], knots[R][2][N])"""
        ladder, knots = np.zeros((self.R, self.N)), np.zeros((self.R, 2, self.N))
        self._check(self.L.blangpt_adapt_preview(self.h, _d(ladder), _d(knots)))
        return ladder, knots

    def reset_round(self):
        self._check(self.L.blangpt_reset_round(self.h))

    def counters(self):
        out = (C.c_uint64 * 4)()
        self._check(self.L.blangpt_counters(self.h, out))
        return dict(evals=int(out[0]), scans=int(out[1]), launches=int(out[2]), executions=int(out[3]))

    def perform_inference(self, nScans, adaptFraction, thinning=1, want=("energy", "swapIndicators", "samples")):
        rs = rounds(nScans, adaptFraction if self.N > 1 else 0.0)
        total = sum(r[0] for r in rs)
        so, bufs = self._outputs(total, want)
        R = self.R
        summ = InferenceSummary()
        keep = {k: np.zeros((64, R)) for k in ("Lambda", "logZ", "minAccept", "meanAccept")}
        rt = np.zeros((64, R), dtype=np.int64)
        for k, v in keep.items():
            setattr(summ, k, _d(v))
        summ.roundTrips = rt.ctypes.data_as(_lp)
        self._check(self.L.blangpt_perform_inference(self.h, nScans, adaptFraction, thinning,
                                                     C.byref(so) if want else None, C.byref(summ)))
        k = summ.nRounds
        out_rounds = [dict(nScans=int(summ.roundNScans[i]), isAdapt=bool(summ.roundIsAdapt[i]),
                           Lambda=keep["Lambda"][i].copy(), logZ=keep["logZ"][i].copy(), roundTrips=rt[i].copy(),
                           minAccept=keep["minAccept"][i].copy(), meanAccept=keep["meanAccept"][i].copy(),
                           millis=summ.roundMillis[i]) for i in range(k)]
        bufs["rounds"] = out_rounds
        return bufs


class PT:
    """Python mirror of blang.engines.internals.factories.PT (option names and defaults kept).

    `nThreads` is ignored.  `targetAccept` resizes the ladder at the last adaptive round (PT.java:156-163) by
    re-creating the engine with the new number of chains, all started from the beta = 1 state.
    """

    def __init__(self, nChains=8, nScans=1000, adaptFraction=0.5, nPassesPerScan=3.0, thinning=1, random=1,
                 usePriorSamples=True, reversible=False, initialization="SCM", targetAccept=None,
                 logNormalizationEstimator="steppingStone", nThreads=None, ladder="EquallySpaced",
                 annealSupport=True, treatNaNAsNegativeInfinity=False, nReplicas=1, device=0, ctasPerChain=0):
        if targetAccept is not None and not (0.0 <= targetAccept < 1.0):
            raise BlangPTError("targetAccept must be in [0,1)")
        self.targetAccept = targetAccept
        self._engine_kw = dict(random=random, nPassesPerScan=nPassesPerScan, usePriorSamples=usePriorSamples,
                               reversible=reversible, annealSupport=annealSupport,
                               treatNaNAsNegativeInfinity=treatNaNAsNegativeInfinity, device=device, ctasPerChain=ctasPerChain)
        if logNormalizationEstimator not in ("steppingStone", "thermodynamicIntegration"):
            raise BlangPTError("unknown logNormalizationEstimator")
        self.nChains, self.nScans, self.adaptFraction = nChains, nScans, adaptFraction
        self.nPassesPerScan, self.thinning, self.random = nPassesPerScan, thinning, random
        self.initialization, self.ladder = initialization, ladder
        self.logNormalizationEstimator = logNormalizationEstimator
        self.engine = Engine(nChains=nChains, nReplicas=nReplicas, random=random, nPassesPerScan=nPassesPerScan,
                             usePriorSamples=usePriorSamples, reversible=reversible, annealSupport=annealSupport,
                             treatNaNAsNegativeInfinity=treatNaNAsNegativeInfinity, device=device,
                             ctasPerChain=ctasPerChain)

    def setSampledModel(self, spec):
        """PT.setSampledModel (PT.java:270-281): lower the model, ladder, informed initialisation."""
        self._spec = spec
        self.engine.set_model(spec)
        if not isinstance(self.ladder, str):
            self.engine.set_ladder(self.ladder)
        elif self.ladder != "EquallySpaced":
            self.engine.set_ladder(make_ladder(self.ladder, self.nChains))
        self.engine.initialize(self.initialization)

    def performInference(self, want=("energy", "swapIndicators", "samples"), results=None, compressed=False):
        """PT.performInference (PT.java:85-113).  With `results` (a folder) the reference's samples/ and
        monitoring/ tables are written there, one batched write per round (blangsdk_b200/results.py);
        `compressed` writes them as .csv.gz (experimentConfigs.tabularWriter.compressed)."""
        if results is None and self.targetAccept is None:
            res = self.engine.perform_inference(self.nScans, self.adaptFraction, self.thinning, want)
            res["finalStats"] = self.engine.round_stats()
            res["ladder"] = self.engine.get_ladder()
            res["counters"] = self.engine.counters()
            return res
        if results is None:
            return self._perform_with_target_accept(want)
        import time
        from .results import PTResultsWriter
        eng = self.engine
        if eng.R != 1:
            raise BlangPTError("result tables are written for one replica at a time")
        w = PTResultsWriter(results, eng.family, eng.N, self._spec["hyper"], thinning=self.thinning,
                            logNormalizationEstimator=self.logNormalizationEstimator, compressed=compressed)
        rs = rounds(self.nScans, self.adaptFraction if eng.N > 1 else 0.0)
        scan, summary = 0, []
        full = ("energy", "logDensity", "nOutOfSupport", "swapIndicators", "samples")
        for r, (n, is_adapt) in enumerate(rs):
            t0 = time.perf_counter()
            out = eng.run_scans(n, want=full)
            w.write_scans(scan, out)
            scan += n
            millis = 1e3 * (time.perf_counter() - t0)
            if eng.N > 1:
                stats, diag = eng.round_stats(), eng.diagnostics()
                w.write_round(r, is_adapt, n, eng.get_ladder()[0], stats, diag, eng.n_samplers, self.nPassesPerScan, millis)
                summary.append(dict(nScans=n, isAdapt=is_adapt, **{k: v for k, v in diag.items() if np.isscalar(v)}))
                if r + 1 < len(rs):
                    eng.adapt()
        w.close()
        return dict(rounds=summary, ladder=eng.get_ladder(), counters=eng.counters(), finalStats=eng.round_stats())

    def _resize_to_target_accept(self):
        """PT.adapt's targetAccept branch (PT.java:156-163): new ladder size from Lambda / (1 - targetAccept),
        every chain re-initialised as a copy of the beta = 1 state."""
        old = self.engine
        ladder = old.target_accept_ladder(self.targetAccept)
        reals, ints = old.get_state(0)
        old.close()
        self.nChains = len(ladder)
        self.engine = Engine(nChains=self.nChains, nReplicas=1, **self._engine_kw)
        self.engine.set_model(self._spec)
        for c in range(self.nChains):
            self.engine.set_state(c, reals if len(reals) else None, ints if len(ints) else None)
        self.engine.set_ladder(ladder)
        self.engine.initialize("COPIES")

    def _perform_with_target_accept(self, want):
        rs = rounds(self.nScans, self.adaptFraction if self.engine.N > 1 else 0.0)
        outs, summary = [], []
        for r, (n, is_adapt) in enumerate(rs):
            out = self.engine.run_scans(n, want=want)
            outs.append(out)
            if self.engine.N > 1:
                d = self.engine.diagnostics()
                summary.append(dict(nScans=n, isAdapt=is_adapt, nChains=self.engine.N,
                                    **{k: v for k, v in d.items() if np.isscalar(v)}))
                if r == len(rs) - 2:
                    self._resize_to_target_accept()
                elif r + 1 < len(rs):
                    self.engine.adapt()
        return dict(rounds=summary, outputs=outs, ladder=self.engine.get_ladder(), counters=self.engine.counters(),
                    finalStats=self.engine.round_stats())
