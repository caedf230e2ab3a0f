package blang.engines.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

/**
 * Host buffers of one batch of scans (struct blangpt_scan_outputs): energies, annealed log-densities, out-of-support
 * counts and swap indicators per (scan, chain position), and the beta = 1 sample per scan.  Layouts as blangpt.h:
 * [scan][replica][position] and [scan][replica][variable]; this shim runs one replica.
 *
 * Sized by the CALLER for the number of scans it asks for; blangpt_perform_inference writes the SUM of
 * blangpt_rounds() scans, which can exceed nScans by one (PT.rounds, PT.java:239-265: 1000 scans at
 * adaptFraction 0.5 -> 1001), so GpuPT batches per round instead.
 *
 * UNCOMPILED SOURCE (no JDK in the build image).
 */
final class ScanOutputs {
  final MemorySegment struct, energy, logDensity, nOutOfSupport, swapIndicators, samples, intSamples;
  final int nScans, nChains, nReals, nInts;

  ScanOutputs(Arena arena, int nScans, int nChains, int nReals, int nInts) {
    this.nScans = nScans; this.nChains = nChains; this.nReals = nReals; this.nInts = nInts;
    long cells = (long) nScans * nChains;
    energy = arena.allocate(JAVA_DOUBLE, cells);
    logDensity = arena.allocate(JAVA_DOUBLE, cells);
    nOutOfSupport = arena.allocate(JAVA_INT, cells);
    swapIndicators = arena.allocate(JAVA_INT, cells);
    samples = nReals > 0 ? arena.allocate(JAVA_DOUBLE, (long) nScans * nReals) : MemorySegment.NULL;
    intSamples = nInts > 0 ? arena.allocate(JAVA_INT, (long) nScans * nInts) : MemorySegment.NULL;
    struct = arena.allocate(Native.SCAN_OUTPUTS);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "energy"), energy);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "logDensity"), logDensity);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "nOutOfSupport"), nOutOfSupport);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "swapIndicators"), swapIndicators);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "samples"), samples);
    struct.set(ADDRESS, Native.offset(Native.SCAN_OUTPUTS, "intSamples"), intSamples);
  }
  double energy(int scan, int chain) { return energy.getAtIndex(JAVA_DOUBLE, (long) scan * nChains + chain); }
  double logDensity(int scan, int chain) { return logDensity.getAtIndex(JAVA_DOUBLE, (long) scan * nChains + chain); }
  int swapIndicator(int scan, int chain) { return swapIndicators.getAtIndex(JAVA_INT, (long) scan * nChains + chain); }
  double sample(int scan, int variable) { return samples.getAtIndex(JAVA_DOUBLE, (long) scan * nReals + variable); }
  int intSample(int scan, int variable) { return intSamples.getAtIndex(JAVA_INT, (long) scan * nInts + variable); }
}
