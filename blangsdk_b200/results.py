"""Output tables in the reference's on-disk format.

The reference's PT engine writes tidy CSV tables every scan through `ExperimentResults.getTabularWriter`
(R/engines/internals/factories/PT.java:116-143,173-229,360-385,387-505,517-537; format:
R/runtime/internals/doc/contents/InputOutput.xtend:185-240).  Here they are written once per ROUND from the
batched device->host outputs of `blangpt_run_scans`, same folders, file names and columns, so that the
reference's consumers (DefaultPostProcessor, ptanalysis/Paths, ladders/FromAnotherExec, user scripts) read them
unchanged:

  samples/     energy, allLogDensities, nOutOfSupport (sample,chain,value); logDensity (sample,value);
               one table per latent variable of the family (index_0,sample,value / sample,value)
  monitoring/  swapIndicators (sample,chain,value); swapStatistics, annealingParameters (isAdapt,round,chain,value);
               swapSummaries (round,lowest,average); energyExplCorrelation (round,chain,beta,isAdapt,value);
               energyExplCorrelationSummaries (round,average,highest); globalLambda (round,value);
               timeToFirstRestart (round,time,effectiveNScans); actualTemperedRestarts, asymptoticRoundTripBound,
               nonAsymptoticRountTrip [sic] (round,count,rate); logNormalizationConstantProgress (round,value);
               cumulativeLambda, lambdaInstantaneous (round,isAdapt,beta,value);
               roundTimings (round,isAdapt,nExplorationSteps,value); runningTimeSummary.tsv
  logNormalizationEstimate.csv (estimator,value)
"""
import gzip
from decimal import Decimal
import os

import numpy as np


def open_table(path, mode="rt"):
    """a tidy CSV as text, gzip-compressed when its name ends in .gz (the reference reads both: BriefIO.readLines,
    DefaultPostProcessor.xtend:107, FromAnotherExec.java:16)"""
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode.replace("t", ""))

# latent variables of each catalogue family as a Blang model would name and shape them
VARIABLES = {
    "normal": [("mu", None, 0, 1), ("sigma2", None, 1, 1)],
    "logistic": [("coefficients", "index_0", 0, None)],
    "mixture": [("means", "index_0", 0, "K"), ("variances", "index_0", "K", "K"), ("pi", "entry", "2K", "K")],
    "betabinomial": [("alpha", None, 0, 1), ("beta", None, 1, 1)],
    "ising": [("vertices", "index_0", 0, None)],
    "linreg": [("coefficients", "index_0", 0, "P"), ("sigma", None, "P", 1)],
    "rockets_latent": [("a", None, 0, 1), ("b", None, 1, 1), ("failureProbabilities", "rocketType", 2, "H"),
                       ("numberOfLaunches", "rocketType", "H2", "H")],
    "hierarchical": [("a", None, 0, 1), ("b", None, 1, 1), ("failureProbabilities", "rocketType", 2, None)],
}
# family "iid": the latent parameters of each observation law, named as in R/distributions/<Name>.bl
IID_VARIABLES = {
    "poisson": ["mean"], "binomial": ["probabilityOfSuccess"], "negbinomial": ["r", "p"],
    "studentt": ["nu", "mu", "sigma"], "geometric": ["p"], "laplace": ["location", "scale"],
    "exponential": ["rate"], "gamma": ["shape", "rate"], "weibull": ["scale", "shape"], "gumbel": ["location", "scale"],
    "logisticdist": ["location", "scale"], "halfstudentt": ["nu", "sigma"], "loglogistic": ["scale", "shape"],
    "gompertz": ["shape", "scale"], "yulesimon": ["rho"], "uniformmax": ["max"],
}


def java_double_to_string(x):
    """java.lang.Double.toString: shortest digits that round-trip; plain notation for 1e-3 <= |x| < 1e7, otherwise
    computerised scientific notation d.dddE[-]n; always at least one digit after the point"""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = "".join(map(str, digits)).rstrip("0") or "0"
    exp10 = len("".join(map(str, Decimal(repr(x)).as_tuple().digits))) + exp - 1     # exponent of the leading digit
    neg = "-" if sign else ""
    if 1e-3 <= abs(x) < 1e7:
        if exp10 >= 0:
            whole, frac = digits[:exp10 + 1].ljust(exp10 + 1, "0"), digits[exp10 + 1:]
        else:
            whole, frac = "0", "0" * (-exp10 - 1) + digits
        return neg + whole + "." + (frac or "0")
    return neg + digits[0] + "." + (digits[1:] or "0") + "E" + str(exp10)


def _fmt(v):
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        v = float(v)
        if 1e-3 <= abs(v) < 1e7:              # Python's shortest repr is Double.toString's plain notation in this range
            return repr(v)
        return java_double_to_string(v)       # NaN, +-Infinity, 0.0, and d.dddE[-]n outside it
    return str(v)


class TabularWriter:
    """One tidy CSV: the header is fixed by the first write (blang.inits TabularWriter behaviour)."""

    def __init__(self, path):
        self.path = path                      # "....csv" or "....csv.gz" (experimentConfigs.tabularWriter.compressed)
        self.columns = None
        self.f = None

    def _open(self):
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        self.f = open_table(self.path, "wt")

    def write(self, *pairs):
        cols = [k for k, _ in pairs]
        if self.f is None:
            self._open()
            self.columns = cols
            self.f.write(",".join(cols) + "\n")
        elif cols != self.columns:
            raise ValueError("columns changed in %s: %s vs %s" % (self.path, cols, self.columns))
        self.f.write(",".join(_fmt(v) for _, v in pairs) + "\n")

    def write_block(self, columns, arrays):
        """vectorised: arrays are equally long 1-D arrays, one per column"""
        if self.f is None:
            self._open()
            self.columns = list(columns)
            self.f.write(",".join(columns) + "\n")
        elif list(columns) != self.columns:
            raise ValueError("columns changed in %s" % self.path)
        rows = zip(*[[_fmt(x) for x in a] for a in arrays])
        self.f.write("".join(",".join(r) + "\n" for r in rows))

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


class ExperimentResults:
    """results folder with lazily created tabular writers (blang.inits ExperimentResults)."""

    def __init__(self, folder, compressed=False):
        self.folder = folder
        self.ext = ".csv.gz" if compressed else ".csv"
        os.makedirs(folder, exist_ok=True)
        self.writers = {}

    def writer(self, sub, name):
        key = (sub, name)
        if key not in self.writers:
            base = os.path.join(self.folder, sub) if sub else self.folder
            self.writers[key] = TabularWriter(os.path.join(base, name + self.ext))
        return self.writers[key]

    def flush_all(self):
        if self.ext.endswith(".gz"):          # "When using .csv.gz, we cannot flush part-way through" (PT.java:432)
            return
        for w in self.writers.values():
            if w.f is not None:
                w.f.flush()

    def close_all(self):
        for w in self.writers.values():
            w.close()


class PTResultsWriter:
    """Writes what PT.performInference writes, one round at a time."""

    def __init__(self, folder, family, n_chains, hyper=None, statisticRecordedMaxChainIndex=100, thinning=1,
                 logNormalizationEstimator="steppingStone", compressed=False):
        self.results = ExperimentResults(folder, compressed)
        self.family, self.N = family, n_chains
        self.max_idx = statisticRecordedMaxChainIndex
        self.thinning = thinning
        self.estimator = logNormalizationEstimator
        self.K = int(hyper["K"]) if hyper and "K" in hyper else None
        self.variables = VARIABLES.get(family)
        if family == "iid":
            self.variables = [(name, None, j, 1) for j, name in enumerate(IID_VARIABLES[hyper["dist"]])]

    def _resolve(self, v, n_reals):
        if v is None:
            return n_reals
        if isinstance(v, str):
            if v == "P":                      # linreg: the coefficients come first, sigma last
                return n_reals - 1
            if v in ("H", "H2"):              # rockets_latent: a, b, then G probabilities, then G counts
                G = (n_reals - 2) // 2
                return G if v == "H" else G + 2
            return {"K": self.K, "2K": 2 * self.K}[v]
        return v

    # ---- per-scan tables (PT.java:116-143, 360-385), written per round from the batched outputs ----
    def write_scans(self, first_scan, out):
        n = out["energy"].shape[0]
        N = self.N
        m = min(N, self.max_idx)
        scans = np.arange(first_scan, first_scan + n)
        samp = np.repeat(scans, m)
        chain = np.tile(np.arange(m), n)
        S = self.results.writer
        if "logDensity" in out:
            S("samples", "allLogDensities").write_block(["sample", "chain", "value"], [samp, chain, out["logDensity"][:, 0, :m].ravel()])
        S("samples", "energy").write_block(["sample", "chain", "value"], [samp, chain, out["energy"][:, 0, :m].ravel()])
        if "nOutOfSupport" in out:
            S("samples", "nOutOfSupport").write_block(["sample", "chain", "value"], [samp, chain, out["nOutOfSupport"][:, 0, :m].ravel()])
        if "logDensity" in out and self.max_idx >= 0:
            S("samples", "logDensity").write_block(["sample", "value"], [scans, out["logDensity"][:, 0, 0]])
        keep = (scans % self.thinning) == 0
        if "samples" in out or "intSamples" in out:
            vals = out["samples"][:, 0, :] if "samples" in out else out["intSamples"][:, 0, :]
            n_reals = vals.shape[1]
            for name, index_col, start, count in self.variables:
                a = self._resolve(start, n_reals)
                c = self._resolve(count, n_reals) if count is not None else n_reals - a
                block = vals[keep, a:a + c]
                ks = scans[keep]
                if index_col is None:
                    S("samples", name).write_block(["sample", "value"], [ks, block[:, 0]])
                else:
                    S("samples", name).write_block([index_col, "sample", "value"],
                                                   [np.tile(np.arange(c), len(ks)), np.repeat(ks, c), block.ravel()])
        if N > 1 and "swapIndicators" in out:
            S("monitoring", "swapIndicators").write_block(
                ["sample", "chain", "value"], [np.repeat(scans, N), np.tile(np.arange(N), n), out["swapIndicators"][:, 0, :].ravel()])

    # ---- per-round tables (PT.java:173-229, 387-505) ----
    def write_round(self, round_index, is_adapt, n_scans, ladder_desc, stats, diag, n_samplers, n_passes, millis):
        W = lambda name: self.results.writer("monitoring", name)
        N = self.N
        acc = stats["swapAcceptMean"][0]
        for i in range(N - 1):   # reportAcceptanceRatios :173-190
            W("swapStatistics").write(("isAdapt", is_adapt), ("round", round_index), ("chain", i), ("value", acc[i]))
            W("annealingParameters").write(("isAdapt", is_adapt), ("round", round_index), ("chain", i), ("value", ladder_desc[i]))
        W("swapSummaries").write(("round", round_index), ("lowest", diag["swapLowest"]), ("average", diag["swapAverage"]))
        var, cov = stats["energyVariance"][0], stats["energyCovariance"][0]
        for c in range(N):       # :400-417
            with np.errstate(divide="ignore", invalid="ignore"):
                corr = float(np.clip(cov[c] / var[c], -1.0, 1.0)) if var[c] == var[c] else float("nan")
            W("energyExplCorrelation").write(("round", round_index), ("chain", c), ("beta", ladder_desc[c]),
                                             ("isAdapt", is_adapt), ("value", corr))
        W("energyExplCorrelationSummaries").write(("round", round_index), ("average", diag["energyCorrelationAverage"]),
                                                  ("highest", diag["energyCorrelationHighest"]))
        W("globalLambda").write(("round", round_index), ("value", diag["Lambda"]))
        W("timeToFirstRestart").write(("round", round_index), ("time", diag["timeToFirstRestart"]),
                                      ("effectiveNScans", diag["effectiveNScans"]))
        W("actualTemperedRestarts").write(("round", round_index), ("count", diag["temperedRestarts"]),
                                          ("rate", diag["temperedRestartRate"]))
        W("asymptoticRoundTripBound").write(("round", round_index), ("count", diag["asymptoticRoundTripCount"]),
                                            ("rate", diag["asymptoticRoundTripRate"]))
        W("nonAsymptoticRountTrip").write(("round", round_index), ("count", diag["nonAsymptoticRoundTripCount"]),
                                          ("rate", diag["nonAsymptoticRoundTripRate"]))
        logz = diag["logNormalizationSteppingStone"] if self.estimator == "steppingStone" else diag["logNormalizationThermodynamic"]
        if logz == logz:
            W("logNormalizationConstantProgress").write(("round", round_index), ("value", logz))
            if not is_adapt:     # :497-504 final estimate in the SCM format
                w = self.results.writer("", "logNormalizationEstimate")
                w.write(("estimator", self.estimator), ("value", logz))
        for i in range(1, 100):  # reportLambdaFunctions :192-210
            beta = i / 100.0
            W("cumulativeLambda").write(("round", round_index), ("isAdapt", is_adapt), ("beta", beta),
                                        ("value", diag["cumulativeLambda"][i - 1]))
            W("lambdaInstantaneous").write(("round", round_index), ("isAdapt", is_adapt), ("beta", beta),
                                           ("value", diag["lambdaInstantaneous"][i - 1]))
        n_expl = n_scans * int(np.floor(n_passes * n_samplers)) * N      # reportRoundTiming :212-229
        W("roundTimings").write(("round", round_index), ("isAdapt", is_adapt), ("nExplorationSteps", n_expl),
                                ("value", int(round(millis))))
        if not is_adapt:
            os.makedirs(os.path.join(self.results.folder, "monitoring"), exist_ok=True)
            with open(os.path.join(self.results.folder, "monitoring", "runningTimeSummary.tsv"), "a") as f:
                f.write("postAdaptTime_ms\t%d\n" % int(round(millis)))
        self.results.flush_all()

    def close(self):
        self.results.close_all()


def replay_round_trips(swap_indicators_csv, start_scan, end_scan):
    """ptanalysis/Paths.xtend:83-117 + nRejuvenations :26-50 replayed from swapIndicators.csv"""
    import csv
    rows = {}
    n_chains = 0
    with open_table(swap_indicators_csv) as f:
        for r in csv.DictReader(f):
            s, c, v = int(r["sample"]), int(r["chain"]), int(r["value"])
            n_chains = max(n_chains, c + 1)
            if start_scan <= s < end_scan:
                rows.setdefault(s, {})[c] = v
    paths = [[c] for c in range(n_chains)]
    for s in sorted(rows):
        just = False
        for c in range(n_chains):
            if just:
                just = False
                continue
            if rows[s][c] == 1:
                paths[c], paths[c + 1] = paths[c + 1], paths[c]
                paths[c].append(c)
                paths[c + 1].append(c + 1)
                just = True
            else:
                paths[c].append(c)
    count = 0
    for p in paths:
        new = False
        for st in p:
            if st == 0 and new:
                new = False
                count += 1
            elif st == n_chains - 1:
                new = True
    return count
