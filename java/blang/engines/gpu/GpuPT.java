package blang.engines.gpu;

import static blang.runtime.Runner.sampleColumn;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.List;
import java.util.Optional;

import org.eclipse.xtext.xbase.lib.Pair;

import bayonet.distributions.Random;
import blang.engines.internals.PosteriorInferenceEngine;
import blang.engines.internals.factories.PT;
import blang.engines.internals.factories.PT.Column;
import blang.engines.internals.factories.PT.InitType;
import blang.engines.internals.factories.PT.LogNormalizationEstimator;
import blang.engines.internals.factories.PT.MonitoringOutput;
import blang.engines.internals.factories.PT.SampleOutput;
import blang.engines.internals.ladders.EquallySpaced;
import blang.engines.internals.ladders.TemperatureLadder;
import blang.inits.Arg;
import blang.inits.DefaultValue;
import blang.inits.GlobalArg;
import blang.inits.experiments.ExperimentResults;
import blang.inits.experiments.tabwriters.TabularWriter;
import blang.inits.experiments.tabwriters.TidySerializer;
import blang.io.BlangTidySerializer;
import blang.runtime.Runner;
import blang.runtime.SampledModel;
import blang.runtime.internals.objectgraph.GraphAnalysis;
import blang.system.Cores;

/**
 * Non-reversible parallel tempering on a B200 through libblangpt.so: a drop-in for blang.engines.internals.factories.PT
 * behind the reference's own plug-in point, PosteriorInferenceEngine (PosteriorInferenceEngine.java:16-22).  Select it
 * with `--engine blang.engines.gpu.GpuPT` (a fully-qualified class name is accepted, InferenceAndRuntime.xtend:415-417);
 * every `--engine.*` option of PT / ParallelTempering keeps its name, default and meaning (ParallelTempering.java:25-40,
 * PT.java:47-82).  Nothing else of the Java stack changes: DSL compilation, GraphAnalysis, Runner, post-processing.
 *
 * Call order, as Runner drives it (Runner.xtend:195,224,238-240): check(ga) -> setSampledModel(sm) -> performInference().
 *   check            recognises the model family (ModelRecogniser) or throws: there is no CPU fallback.
 *   setSampledModel  creates the handle, uploads the data, sets the ladder, initialises the chains (COPIES / FORWARD /
 *                    SCM, PT.java:284-329), and verifies the lowering: the engine's annealed log-density of the initial
 *                    state must equal SampledModel.logDensity() at two exponents to 1e-9 (north_star's parity bar).
 *   performInference the round loop of PT.java:85-113 with the scans of a round batched into one native call; the CSV
 *                    tables are written with the reference's own writers, same names and columns (PT.java:116-143,
 *                    173-229, 360-385, 387-505).
 *
 * UNCOMPILED SOURCE: the build image has no JDK (DESIGN.md section 1); the same C ABI is exercised end to end by the
 * Python harness (blangsdk_b200/engine.py) and by tests/c_abi_driver.c.
 */
public class GpuPT implements PosteriorInferenceEngine {
  @GlobalArg public ExperimentResults results = new ExperimentResults();

  // ---- ParallelTempering.java:25-40 ----
  @Arg(description = "Ignored: the chains run on the GPU; results never depend on it (InferenceAndRuntime.xtend:398-402)")
  @DefaultValue("Dynamic") public Cores nThreads = Cores.dynamic();
  @Arg(description = "The annealing schedule to use or if adaptation is used, the initial value")
  @DefaultValue("EquallySpaced") public TemperatureLadder ladder = new EquallySpaced();
  @Arg @DefaultValue("8") public int nChains = 8;
  @Arg @DefaultValue("true") public boolean usePriorSamples = true;
  @Arg @DefaultValue("false") public boolean reversible = false;
  // ---- PT.java:47-82 ----
  @Arg @DefaultValue("1_000") public int nScans = 1_000;
  @Arg(description = "Set to zero for disabling schedule adaptation") @DefaultValue("0.5") public double adaptFraction = 0.5;
  @Arg @DefaultValue("3") public double nPassesPerScan = 3;
  @Arg(description = "Collect statistics every thinning iteration") @DefaultValue("1") public int thinning = 1;
  @Arg @DefaultValue("1") public Random random = new Random(1);
  @Arg public Optional<Double> targetAccept = Optional.empty();
  @Arg @DefaultValue("100") public int scmNParticles = 100;                  // PT.scmInit defaults, PT.java:68-72,507-515
  @Arg @DefaultValue("0.9") public double scmThreshold = 0.9;
  @Arg @DefaultValue("0.5") public double scmResamplingESSThreshold = 0.5;
  @Arg @DefaultValue("SCM") public InitType initialization = InitType.SCM;
  @Arg @DefaultValue("steppingStone") public LogNormalizationEstimator logNormalizationEstimator = LogNormalizationEstimator.steppingStone;
  @Arg @DefaultValue("100") public int statisticRecordedMaxChainIndex = 100;
  @Arg(description = "Not supported by the GPU engine (states of the other chains stay on the device)")
  @DefaultValue("false") public boolean storeSamplesForAllChains = false;
  // ---- GPU-specific ----
  @Arg(description = "CUDA device ordinal") @DefaultValue("0") public int device = 0;
  @Arg(description = "Thread-block cluster size per chain; 0 = chosen by the library") @DefaultValue("0") public int ctasPerChain = 0;

  private final Arena arena = Arena.ofShared();
  private MemorySegment handle = MemorySegment.NULL;
  private SampledModel model;
  private LoweredModel lowered;
  private boolean annealSupport = true, treatNaNAsNegativeInfinity = false;   // Runner options, read from GraphAnalysis
  private BlangTidySerializer tidySerializer, densitySerializer, swapIndicatorSerializer;

  /** Runner.preprocess calls this before setSampledModel (Runner.xtend:195): refuse what the GPU cannot run. */
  @Override public void check(GraphAnalysis analysis) {
    if (storeSamplesForAllChains) throw new RuntimeException("GpuPT: storeSamplesForAllChains is not supported");
    if (nChains < 1) throw new RuntimeException("Number of tempering chains must be greater than zero.");
    annealSupport = analysis.annealSupport;                       // Runner.xtend:72-73 via GraphAnalysis.java:61-62
    treatNaNAsNegativeInfinity = analysis.treatNaNAsNegativeInfinity;
    lowered = ModelRecogniser.recognise(analysis);                // throws for models outside the catalogue
  }

  @Override public void setSampledModel(SampledModel model) {
    this.model = model;
    try {
      createHandle(nChains);
      List<Double> betas = ladder.temperingParameters(nChains);   // the reference's own ladders/*.java, index 0 = 1.0
      setLadder(betas);
      if (initialization == InitType.COPIES) copyPrototypeIntoAllChains();
      if (initialization == InitType.SCM)
        Native.check((int) Native.SET_SCM_OPTIONS.invoke(handle, scmNParticles, scmThreshold, scmResamplingESSThreshold), handle);
      verifyLowering();
      Native.check((int) Native.INITIALIZE.invoke(handle, initialization.ordinal()), handle);   // COPIES 0, FORWARD 1, SCM 2
      tidySerializer = new BlangTidySerializer(results.child(Runner.SAMPLES_FOLDER));           // PT.initSerializers :336-345
      densitySerializer = new BlangTidySerializer(results.child(Runner.SAMPLES_FOLDER));
      swapIndicatorSerializer = new BlangTidySerializer(results.child(Runner.MONITORING_FOLDER));
    } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
  }

  private void createHandle(int chains) throws Throwable {
    MemorySegment cfg = arena.allocate(Native.CONFIG);
    Native.DEFAULT_CONFIG.invoke(cfg);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "nChains"), chains);
    cfg.set(JAVA_LONG, Native.offset(Native.CONFIG, "random"), random.nextLong());   // one draw of the engine's stream seeds Philox
    cfg.set(JAVA_DOUBLE, Native.offset(Native.CONFIG, "nPassesPerScan"), nPassesPerScan);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "usePriorSamples"), usePriorSamples ? 1 : 0);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "reversible"), reversible ? 1 : 0);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "annealSupport"), annealSupport ? 1 : 0);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "treatNaNAsNegativeInfinity"), treatNaNAsNegativeInfinity ? 1 : 0);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "device"), device);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "ctasPerChain"), ctasPerChain);
    cfg.set(JAVA_INT, Native.offset(Native.CONFIG, "statisticRecordedMaxChainIndex"), statisticRecordedMaxChainIndex);
    MemorySegment out = arena.allocate(ADDRESS);
    Native.check((int) Native.CREATE.invoke(cfg, out), null);
    handle = out.get(ADDRESS, 0);
    Native.check((int) Native.SET_MODEL.invoke(handle, lowered.familyId, lowered.arraysStruct(arena), lowered.nArrays(),
        lowered.hyperArray(arena), lowered.nHyper()), handle);
    int nr = (int) Native.N_REALS.invoke(handle), ni = (int) Native.N_INTS.invoke(handle);
    if (nr != lowered.nReals() || ni != lowered.nInts())
      throw new RuntimeException("GpuPT: the recogniser found " + lowered.nReals() + " real / " + lowered.nInts()
          + " integer latent variables, family " + lowered.familyName + " has " + nr + " / " + ni);
  }

  private void setLadder(List<Double> betas) throws Throwable {
    double[] b = betas.stream().mapToDouble(Double::doubleValue).toArray();
    Native.check((int) Native.SET_LADDER.invoke(handle, arena.allocateFrom(JAVA_DOUBLE, b), (long) b.length), handle);
  }

  private void copyPrototypeIntoAllChains() throws Throwable {   // PT.initialize: states[i] = prototype.duplicate()
    MemorySegment reals = lowered.currentReals(arena), ints = lowered.currentInts(arena);
    for (int c = 0; c < nChains; c++) Native.check((int) Native.SET_STATE.invoke(handle, 0L, c, reals, ints), handle);
  }

  /**
   * The lowering is only trusted after a numerical cross-check against the Java model itself: with every chain at the
   * prototype state, the engine's loglik / fixed sums must reproduce SampledModel.logDensity(beta) at beta = 1 and at an
   * interior exponent (SampledModel.java:161-176,190-197).  A recogniser bug therefore fails loudly before any sampling.
   */
  private void verifyLowering() throws Throwable {
    copyPrototypeIntoAllChains();
    MemorySegment ld = arena.allocate(JAVA_DOUBLE, nChains), ll = arena.allocate(JAVA_DOUBLE, nChains),
        no = arena.allocate(JAVA_INT, nChains), fx = arena.allocate(JAVA_DOUBLE, nChains);
    Native.check((int) Native.LOG_DENSITY.invoke(handle, ld, ll, no, fx), handle);
    double loglik = ll.getAtIndex(JAVA_DOUBLE, 0), fixed = fx.getAtIndex(JAVA_DOUBLE, 0);
    int noos = no.getAtIndex(JAVA_INT, 0);
    for (double beta : new double[] {1.0, 0.37}) {
      double java = model.logDensity(beta);
      double gpu = noos == 0 ? fixed + beta * loglik
          : (!annealSupport || beta >= 1.0 ? Double.NEGATIVE_INFINITY : fixed + beta * loglik + noos * (-beta * 1E100));
      boolean same = (Double.isInfinite(java) && java == gpu) || Math.abs(java - gpu) <= 1e-9 * Math.max(1.0, Math.abs(java));
      if (!same) throw new RuntimeException("GpuPT: lowering check failed for family " + lowered.familyName + " at exponent " + beta
          + ": SampledModel.logDensity = " + java + ", engine = " + gpu + " (the model is not what the recogniser took it for)");
    }
  }

  /** PT.performInference, PT.java:85-113. */
  @Override public void performInference() {
    try {
      MemorySegment rn = arena.allocate(JAVA_INT, 64), ra = arena.allocate(JAVA_INT, 64);
      int nRounds = (int) Native.ROUNDS.invoke(nScans, nChains == 1 ? 0.0 : adaptFraction, rn, ra);   // PT.rounds :239-265, :235
      if (nRounds < 0) throw new RuntimeException("GpuPT: adaptFraction must be in [0, 1)");
      int scan = 0;
      for (int r = 0; r < nRounds; r++) {
        final int n = rn.getAtIndex(JAVA_INT, r);
        final boolean isAdapt = ra.getAtIndex(JAVA_INT, r) != 0;
        final long t0 = System.nanoTime();
        try (Arena round = Arena.ofConfined()) {
          ScanOutputs o = new ScanOutputs(round, n, nChains, lowered.nReals(), lowered.nInts());
          Native.check((int) Native.RUN_SCANS.invoke(handle, n, o.struct), handle);      // one native call, one D2H per round
          for (int s = 0; s < n; s++, scan++) {                                          // the per-scan rows of PT.java:93-101
            recordEnergyStatistics(o, s, scan);
            if (scan % thinning == 0) recordSamples(o, s, scan);
            if (nChains > 1) recordSwapIndicators(o, s, scan);
          }
          if (nChains > 1) {   // in the last round this is done for instrumentation only (PT.java:103-107)
            RoundStats st = new RoundStats(round, nChains);
            st.fetch(handle);
            List<Double> betas = currentLadder();
            reportAcceptanceRatios(st, betas, r, isAdapt);
            reportDiagnostics(st, betas, r, isAdapt);
            reportLambdaFunctions(st, r, isAdapt);
            if (r == nRounds - 2 && targetAccept.isPresent()) resizeToTargetAccept();    // PT.adapt(finalAdapt), :156-163
            else if (r + 1 < nRounds) Native.check((int) Native.ADAPT.invoke(handle, MemorySegment.NULL), handle);
            // after the last round the reference still replaces its in-memory ladder; nothing reads it afterwards.
            // blangpt_adapt_preview returns that ladder without touching the accumulators, for callers who want it.
          }
        }
        reportRoundTiming(r, isAdapt, n, (System.nanoTime() - t0) / 1_000_000L);
        results.flushAll();
      }
    } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    finally { close(); }
  }

  private void close() {
    try { if (!handle.equals(MemorySegment.NULL)) Native.DESTROY.invoke(handle); } catch (Throwable ignored) {}
    handle = MemorySegment.NULL;
  }

  private List<Double> currentLadder() throws Throwable {
    MemorySegment b = arena.allocate(JAVA_DOUBLE, nChains);
    Native.check((int) Native.GET_LADDER.invoke(handle, b, (long) nChains), handle);
    Double[] v = new Double[nChains];
    for (int i = 0; i < nChains; i++) v[i] = b.getAtIndex(JAVA_DOUBLE, i);
    return List.of(v);
  }

  /** PT.adapt's targetAccept branch: new ladder size from Lambda, all chains restart from the beta = 1 state. */
  private void resizeToTargetAccept() throws Throwable {
    final int capacity = 1 << 16;
    MemorySegment out = arena.allocate(JAVA_DOUBLE, capacity), nOut = arena.allocate(JAVA_INT);
    Native.check((int) Native.TARGET_ACCEPT_LADDER.invoke(handle, 0L, targetAccept.get().doubleValue(), out, capacity, nOut), handle);
    int newN = nOut.get(JAVA_INT, 0);
    MemorySegment reals = arena.allocate(JAVA_DOUBLE, Math.max(1, lowered.nReals())), ints = arena.allocate(JAVA_INT, Math.max(1, lowered.nInts()));
    Native.check((int) Native.GET_STATE.invoke(handle, 0L, 0, reals, ints), handle);
    close();
    nChains = newN;
    createHandle(newN);
    for (int c = 0; c < newN; c++) Native.check((int) Native.SET_STATE.invoke(handle, 0L, c, reals, ints), handle);
    Native.check((int) Native.SET_LADDER.invoke(handle, out, (long) newN), handle);
    Native.check((int) Native.INITIALIZE.invoke(handle, Native.INIT_COPIES), handle);
  }

  // ---- the reference's tables, same files / columns / order ----
  private TabularWriter writer(MonitoringOutput output) {
    return results.child(Runner.MONITORING_FOLDER).getTabularWriter(output.toString());          // PT.java:517-520
  }

  @SuppressWarnings("unchecked")
  private void recordEnergyStatistics(ScanOutputs o, int s, int scan) {                          // PT.java:116-143
    for (int i = 0; i < Math.min(nChains, statisticRecordedMaxChainIndex); i++) {
      densitySerializer.serialize(o.logDensity(s, i), SampleOutput.allLogDensities.toString(), Pair.of(sampleColumn, scan), Pair.of(Column.chain, i));
      densitySerializer.serialize(o.energy(s, i), SampleOutput.energy.toString(), Pair.of(sampleColumn, scan), Pair.of(Column.chain, i));
      densitySerializer.serialize(o.nOutOfSupport.getAtIndex(JAVA_INT, (long) s * nChains + i), SampleOutput.nOutOfSupport.toString(),
          Pair.of(sampleColumn, scan), Pair.of(Column.chain, i));
    }
    if (statisticRecordedMaxChainIndex >= 0)
      densitySerializer.serialize(o.logDensity(s, 0), SampleOutput.logDensity.toString(), Pair.of(sampleColumn, scan));
  }

  @SuppressWarnings("unchecked")
  private void recordSamples(ScanOutputs o, int s, int scan) {                                   // PT.java:360-372
    lowered.writeBack(o, s);                      // WritableRealVar.set / WritableIntVar.set on the model's own variables
    model.getSampleWriter(tidySerializer).write(Pair.of(sampleColumn, scan));
  }

  @SuppressWarnings("unchecked")
  private void recordSwapIndicators(ScanOutputs o, int s, int scan) {                            // PT.java:374-385
    for (int c = 0; c < nChains; c++)
      swapIndicatorSerializer.serialize(o.swapIndicator(s, c), MonitoringOutput.swapIndicators.toString(),
          Pair.of(sampleColumn, scan), Pair.of(Column.chain, c));
  }

  private void reportAcceptanceRatios(RoundStats st, List<Double> betas, int round, boolean isAdapt) {   // PT.java:173-190
    TabularWriter swap = writer(MonitoringOutput.swapStatistics), ann = writer(MonitoringOutput.annealingParameters);
    for (int i = 0; i < nChains - 1; i++) {
      swap.write(Pair.of(Column.isAdapt, isAdapt), Pair.of(Column.round, round), Pair.of(Column.chain, i), Pair.of(TidySerializer.VALUE, st.swapAccept(i)));
      ann.write(Pair.of(Column.isAdapt, isAdapt), Pair.of(Column.round, round), Pair.of(Column.chain, i), Pair.of(TidySerializer.VALUE, betas.get(i)));
    }
  }

  private void reportDiagnostics(RoundStats st, List<Double> betas, int round, boolean isAdapt) {        // PT.java:387-505
    Pair<?, ?> r = Pair.of(Column.round, round);
    writer(MonitoringOutput.swapSummaries).printAndWrite(r, Pair.of(Column.lowest, st.diag("swapLowest")), Pair.of(Column.average, st.diag("swapAverage")));
    for (int c = 0; c < nChains; c++) {
      double corr = Math.max(-1.0, Math.min(1.0, st.energyCovariance(c) / st.energyVariance(c)));
      writer(MonitoringOutput.energyExplCorrelation).write(r, Pair.of(Column.chain, c), Pair.of(Column.beta, betas.get(c)),
          Pair.of(Column.isAdapt, isAdapt), Pair.of(TidySerializer.VALUE, corr));
    }
    writer(MonitoringOutput.energyExplCorrelationSummaries).printAndWrite(r, Pair.of(Column.average, st.diag("energyCorrelationAverage")),
        Pair.of(Column.highest, st.diag("energyCorrelationHighest")));
    writer(MonitoringOutput.globalLambda).printAndWrite(r, Pair.of(TidySerializer.VALUE, st.diag("Lambda")));
    writer(MonitoringOutput.timeToFirstRestart).printAndWrite(r, Pair.of(Column.time, st.diag("timeToFirstRestart")),
        Pair.of(Column.effectiveNScans, st.diag("effectiveNScans")));
    writer(MonitoringOutput.actualTemperedRestarts).printAndWrite(r, Pair.of(Column.count, st.diagLong("temperedRestarts")),
        Pair.of(Column.rate, st.diag("temperedRestartRate")));
    writer(MonitoringOutput.asymptoticRoundTripBound).printAndWrite(r, Pair.of(Column.count, st.diag("asymptoticRoundTripCount")),
        Pair.of(Column.rate, st.diag("asymptoticRoundTripRate")));
    writer(MonitoringOutput.nonAsymptoticRountTrip).printAndWrite(r, Pair.of(Column.count, st.diag("nonAsymptoticRoundTripCount")),
        Pair.of(Column.rate, st.diag("nonAsymptoticRoundTripRate")));
    double logZ = logNormalizationEstimator == LogNormalizationEstimator.steppingStone
        ? st.diag("logNormalizationSteppingStone") : st.diag("logNormalizationThermodynamic");
    if (!Double.isNaN(logZ)) {
      writer(MonitoringOutput.logNormalizationConstantProgress).printAndWrite(r, Pair.of(TidySerializer.VALUE, logZ));
      if (!isAdapt)   // final estimate in the SCM engine's format, PT.java:497-504
        results.getTabularWriter(Runner.LOG_NORMALIZATION_ESTIMATE).write(
            Pair.of(Runner.LOG_NORMALIZATION_ESTIMATOR, logNormalizationEstimator), Pair.of(TidySerializer.VALUE, logZ));
    }
  }

  private void reportLambdaFunctions(RoundStats st, int round, boolean isAdapt) {                        // PT.java:192-210
    for (int i = 1; i < PT._lamdbaDiscretization; i++) {
      double beta = ((double) i) / ((double) PT._lamdbaDiscretization);
      writer(MonitoringOutput.cumulativeLambda).write(Pair.of(Column.round, round), Pair.of(Column.isAdapt, isAdapt),
          Pair.of(Column.beta, beta), Pair.of(TidySerializer.VALUE, st.cumulativeLambda(i - 1)));
      writer(MonitoringOutput.lambdaInstantaneous).write(Pair.of(Column.round, round), Pair.of(Column.isAdapt, isAdapt),
          Pair.of(Column.beta, beta), Pair.of(TidySerializer.VALUE, st.lambdaInstantaneous(i - 1)));
    }
  }

  private void reportRoundTiming(int round, boolean isAdapt, int nScansInRound, long millis) throws Throwable {   // PT.java:212-229
    int nSamplers = (int) Native.N_SAMPLERS.invoke(handle);
    int perChain = nScansInRound * (int) Math.floor(nPassesPerScan * nSamplers);
    writer(MonitoringOutput.roundTimings).write(Pair.of(Column.round, round), Pair.of(Column.isAdapt, isAdapt),
        Pair.of(Column.nExplorationSteps, perChain * nChains), Pair.of(TidySerializer.VALUE, millis));
    if (!isAdapt) {
      try {
        results.child(Runner.MONITORING_FOLDER).getAutoClosedBufferedWriter(Runner.RUNNING_TIME_SUMMARY)
            .append("postAdaptTime_ms\t" + millis + "\n");
      } catch (Exception e) { /* as the reference: best effort */ }
    }
  }
}
