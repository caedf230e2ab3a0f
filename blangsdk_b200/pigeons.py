"""Pigeons line protocol over the C ABI.

The reference's only cross-process boundary is `factories/Pigeons.java:30-70`: one JVM per replica reads
stdin lines and answers on stdout, so Pigeons.jl can run Blang models under MPI:

    log_potential(<beta>)   ->  response(<logDensity at beta of the current state>)
    call_sampler!(<beta>)   ->  set the exponent; beta == 0: forwardSample, else posteriorSamplingScan(nPassesPerScan)
                                ->  response()

This server speaks the same protocol with the replica state resident on the GPU (a one-chain engine whose
exponent is moved with blangpt_set_ladder).  Usage: python -m blangsdk_b200.pigeons --family normal ...
"""
import re
import sys

from .engine import Engine

_LOG = re.compile(r"^log_potential\((.*)\)$")
_SAMPLE = re.compile(r"^call_sampler!\((.*)\)$")


class PigeonsServer:
    def __init__(self, spec, random=1, nPassesPerScan=3.0, device=0, initialization="FORWARD"):
        self.engine = Engine(nChains=1, random=random, nPassesPerScan=nPassesPerScan, device=device)
        self.engine.set_model(spec)
        self.engine.initialize(initialization)

    def log_potential(self, beta):
        """SampledModel.logDensity(beta) of the current state (Pigeons.java:66-70)"""
        d = self.engine.log_density()
        fixed, loglik, noos = d["fixed"][0, 0], d["loglik"][0, 0], d["noos"][0, 0]
        if noos > 0:
            if beta >= 1.0 or not self.engine.cfg.annealSupport:
                return float("-inf")
            return fixed + beta * loglik + noos * (-beta * 1e100)
        return fixed + beta * loglik

    def call_sampler(self, beta):
        """Pigeons.java:56-64"""
        self.engine.set_ladder([beta])
        self.engine.run_scans(1, want=())

    def handle(self, line):
        line = line.strip()
        m = _LOG.match(line)
        if m:
            return "response(%r)" % float(self.log_potential(float(m.group(1))))
        m = _SAMPLE.match(line)
        if m:
            self.call_sampler(float(m.group(1)))
            return "response()"
        raise ValueError("unknown command: " + line)

    def serve(self, inp=sys.stdin, out=sys.stdout):
        for line in inp:
            if not line.strip():
                continue
            out.write(self.handle(line) + "\n")
            out.flush()


def main(argv=None):
    import argparse
    from . import datasets
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="normal", choices=["normal", "logistic", "mixture", "ising", "betabinomial"])
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--random", type=int, default=1)
    ap.add_argument("--nPassesPerScan", type=float, default=3.0)
    args = ap.parse_args(argv)
    spec = {"normal": lambda: datasets.normal_meanvar(args.n), "logistic": lambda: datasets.logistic_regression(args.n, 8),
            "mixture": lambda: datasets.gaussian_mixture(args.n), "ising": lambda: datasets.ising(16),
            "betabinomial": lambda: datasets.beta_binomial(0)}[args.family]()
    PigeonsServer(spec, random=args.random, nPassesPerScan=args.nPassesPerScan).serve()


if __name__ == "__main__":
    main()
