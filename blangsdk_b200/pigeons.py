"""Pigeons line protocol over the C ABI.

The reference's only cross-process boundary is `factories/Pigeons.java:30-70`: one JVM per replica reads
stdin lines and answers on stdout, so Pigeons.jl can run Blang models under MPI:

    log_potential(<beta>)   ->  response(<logDensity at beta of the current state>)
    call_sampler!(<beta>)   ->  set the exponent; beta == 0: forwardSample, else posteriorSamplingScan(nPassesPerScan)
                                ->  response()

This server speaks the same protocol with the replica state resident on the GPU (a one-chain engine whose
exponent is moved with blangpt_set_ladder).  Usage: python -m blangsdk_b200.pigeons --family normal ...

`BatchedPigeonsServer` holds MANY replicas in one handle (nReplicas = R, one chain each, every replica at its own
exponent) so that one process -- one GPU -- serves a whole Pigeons population and every exploration step is a single
kernel launch over all replicas.  Two ways to talk to it, both built from the reference's two commands:
  * vector form:   log_potential(b_0,...,b_{R-1}) -> response(v_0,...,v_{R-1});  call_sampler!(b_0,...,b_{R-1}) -> response()
  * tagged form:   "@i log_potential(b)" / "@i call_sampler!(b)" for replica i, answered as "@i response(...)".  Pigeons
    explores all replicas in lockstep, so tagged call_sampler! requests are held until every replica has one pending,
    then served by one launch.
Numbers are printed as Java prints them (Double.toString: "-Infinity", "NaN", "1.0E10"), because the other end parses
what `System.out.println("response(" + result + ")")` writes (Pigeons.java:66-70).
"""
import re
import sys
from decimal import Decimal

from .engine import Engine
from .results import java_double_to_string  # noqa: F401  (re-exported: the protocol prints numbers as Java does)


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
            return "response(%s)" % java_double_to_string(self.log_potential(float(m.group(1))))
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


class BatchedPigeonsServer:
    """R replicas in one handle; see the module docstring for the two request forms."""

    def __init__(self, spec, nReplicas, random=1, nPassesPerScan=3.0, device=0, initialization="FORWARD"):
        self.R = int(nReplicas)
        self.engine = Engine(nChains=1, nReplicas=self.R, random=random, nPassesPerScan=nPassesPerScan, device=device)
        self.engine.set_model(spec)
        self.engine.initialize(initialization)
        self.betas = [1.0] * self.R
        self.pending = {}            # replica -> beta of a held call_sampler! request

    def log_potentials(self, betas):
        d = self.engine.log_density()
        out = []
        for r, beta in enumerate(betas):
            fixed, loglik, noos = d["fixed"][r, 0], d["loglik"][r, 0], d["noos"][r, 0]
            if noos > 0:
                out.append(float("-inf") if (beta >= 1.0 or not self.engine.cfg.annealSupport)
                           else fixed + beta * loglik + noos * (-beta * 1e100))
            else:
                out.append(fixed + beta * loglik)
        return out

    def call_samplers(self, betas):
        """every replica explores at its own exponent in ONE launch (beta == 0: forwardSample, Pigeons.java:56-64)"""
        self.betas = [float(b) for b in betas]
        self.engine.set_ladder(self.betas)
        self.engine.run_scans(1, want=())

    def handle(self, line):
        """one request line -> list of response lines (possibly empty while tagged requests are being collected)"""
        line = line.strip()
        tag = None
        m = re.match(r"^@(\d+)\s+(.*)$", line)
        if m:
            tag, line = int(m.group(1)), m.group(2)
            if not 0 <= tag < self.R:
                raise ValueError("replica %d out of range" % tag)
        ml, ms = _LOG.match(line), _SAMPLE.match(line)
        if not (ml or ms):
            raise ValueError("unknown command: " + line)
        args = [float(v) for v in (ml or ms).group(1).split(",")]
        if tag is None:
            if len(args) != self.R:
                raise ValueError("expected %d exponents, one per replica" % self.R)
            if ml:
                return ["response(%s)" % ",".join(java_double_to_string(v) for v in self.log_potentials(args))]
            self.call_samplers(args)
            return ["response()"]
        if len(args) != 1:
            raise ValueError("a tagged request carries one exponent")
        if ml:
            betas = list(self.betas); betas[tag] = args[0]
            return ["@%d response(%s)" % (tag, java_double_to_string(self.log_potentials(betas)[tag]))]
        self.pending[tag] = args[0]
        if len(self.pending) < self.R:
            return []
        betas = [self.pending[r] for r in range(self.R)]
        self.pending = {}
        self.call_samplers(betas)
        return ["@%d response()" % r for r in range(self.R)]

    def serve(self, inp=sys.stdin, out=sys.stdout):
        for line in inp:
            if not line.strip():
                continue
            for resp in self.handle(line):
                out.write(resp + "\n")
            out.flush()


def main(argv=None):
    import argparse
    from . import datasets
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="normal", choices=["normal", "logistic", "mixture", "ising", "betabinomial"])
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--random", type=int, default=1)
    ap.add_argument("--nPassesPerScan", type=float, default=3.0)
    ap.add_argument("--replicas", type=int, default=1, help="> 1: batched server, all replicas in one handle")
    args = ap.parse_args(argv)
    spec = {"normal": lambda: datasets.normal_meanvar(args.n), "logistic": lambda: datasets.logistic_regression(args.n, 8),
            "mixture": lambda: datasets.gaussian_mixture(args.n), "ising": lambda: datasets.ising(16),
            "betabinomial": lambda: datasets.beta_binomial(0)}[args.family]()
    if args.replicas > 1:
        BatchedPigeonsServer(spec, args.replicas, random=args.random, nPassesPerScan=args.nPassesPerScan).serve()
    else:
        PigeonsServer(spec, random=args.random, nPassesPerScan=args.nPassesPerScan).serve()


if __name__ == "__main__":
    main()
