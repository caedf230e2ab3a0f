package blang.engines.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

/**
 * One round's accumulators (struct blangpt_round_stats) and derived diagnostics (struct blangpt_diagnostics), i.e.
 * everything PT.reportAcceptanceRatios (PT.java:173-190), reportParallelTemperingDiagnostics (:387-505) and
 * reportLambdaFunctions (:192-210) print, computed by the library from its device accumulators.
 *
 * UNCOMPILED SOURCE (no JDK in the build image).
 */
final class RoundStats {
  final MemorySegment struct, diagnostics;
  final MemorySegment swapAcceptMean, energyMean, energyVariance, energyCovariance, logSum, logSumN, roundTrips, Lambda,
      logZSteppingStone, logZThermodynamic;
  final int nChains;

  RoundStats(Arena arena, int nChains) {
    this.nChains = nChains;
    swapAcceptMean = arena.allocate(JAVA_DOUBLE, nChains);
    energyMean = arena.allocate(JAVA_DOUBLE, nChains);
    energyVariance = arena.allocate(JAVA_DOUBLE, nChains);
    energyCovariance = arena.allocate(JAVA_DOUBLE, nChains);
    logSum = arena.allocate(JAVA_DOUBLE, nChains);
    logSumN = arena.allocate(JAVA_LONG, nChains);
    roundTrips = arena.allocate(JAVA_LONG, 1);
    Lambda = arena.allocate(JAVA_DOUBLE, 1);
    logZSteppingStone = arena.allocate(JAVA_DOUBLE, 1);
    logZThermodynamic = arena.allocate(JAVA_DOUBLE, 1);
    struct = arena.allocate(Native.ROUND_STATS);
    String[] names = {"swapAcceptMean", "energyMean", "energyVariance", "energyCovariance", "logSum", "logSumN", "roundTrips",
        "Lambda", "logZSteppingStone", "logZThermodynamic"};
    MemorySegment[] segs = {swapAcceptMean, energyMean, energyVariance, energyCovariance, logSum, logSumN, roundTrips, Lambda,
        logZSteppingStone, logZThermodynamic};
    for (int k = 0; k < names.length; k++) struct.set(ADDRESS, Native.offset(Native.ROUND_STATS, names[k]), segs[k]);
    diagnostics = arena.allocate(Native.DIAGNOSTICS);
  }
  /** read both structs for the round that just ended */
  void fetch(MemorySegment handle) throws Throwable {
    Native.check((int) Native.GET_ROUND_STATS.invoke(handle, struct), handle);
    Native.check((int) Native.GET_DIAGNOSTICS.invoke(handle, 0L, diagnostics), handle);
  }
  double swapAccept(int chain) { return swapAcceptMean.getAtIndex(JAVA_DOUBLE, chain); }
  double energyMean(int chain) { return energyMean.getAtIndex(JAVA_DOUBLE, chain); }
  double energyVariance(int chain) { return energyVariance.getAtIndex(JAVA_DOUBLE, chain); }
  double energyCovariance(int chain) { return energyCovariance.getAtIndex(JAVA_DOUBLE, chain); }
  long roundTrips() { return roundTrips.getAtIndex(JAVA_LONG, 0); }
  double diag(String field) { return diagnostics.get(JAVA_DOUBLE, Native.offset(Native.DIAGNOSTICS, field)); }
  long diagLong(String field) { return diagnostics.get(JAVA_LONG, Native.offset(Native.DIAGNOSTICS, field)); }
  double cumulativeLambda(int i) { return diagnostics.get(JAVA_DOUBLE, Native.offset(Native.DIAGNOSTICS, "cumulativeLambda") + 8L * i); }
  double lambdaInstantaneous(int i) { return diagnostics.get(JAVA_DOUBLE, Native.offset(Native.DIAGNOSTICS, "lambdaInstantaneous") + 8L * i); }
}
