package blang.engines.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemoryLayout.PathElement;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

/**
 * Panama (java.lang.foreign, Java 22) bindings of include/blangpt.h: one downcall handle per exported symbol, and the
 * struct layouts in the header's field order.  Nothing here is specific to a model; GpuPT is the only caller.
 *
 * UNCOMPILED SOURCE: the build image has no JDK (DESIGN.md section 1).  A JNI stub would bind the same symbols.
 * The library is looked up from -Dblangpt.lib=/path/to/libblangpt.so (default: libblangpt.so on java.library.path).
 */
final class Native {
  private Native() {}

  static final Linker LINKER = Linker.nativeLinker();
  static final SymbolLookup LIB =
      SymbolLookup.libraryLookup(System.getProperty("blangpt.lib", "libblangpt.so"), Arena.global());

  private static MethodHandle fn(String name, FunctionDescriptor fd) {
    return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
  }
  private static FunctionDescriptor i(MemoryLayout... args) { return FunctionDescriptor.of(JAVA_INT, args); }

  // status codes of blangpt.h
  static final int OK = 0, EINVAL = 1, ECUDA = 2, ESTATE = 3, ENUMERIC = 4, EUNSUPPORTED = 5;
  // model families (blangpt.h BLANGPT_FAM_*)
  static final int FAM_NORMAL = 1, FAM_LOGISTIC = 2, FAM_MIXTURE = 3, FAM_ISING = 4, FAM_BETABINOMIAL = 5, FAM_IID = 6,
      FAM_HIERARCHICAL = 7, FAM_LINREG = 8, FAM_ROCKETS_LATENT = 9, FAM_MIXTURE_INDICATORS = 10, FAM_CHANGEPOINT = 11,
      FAM_ANNEAL_PAIR = 12;
  static final int F64 = 0, I32 = 1;
  static final int INIT_COPIES = 0, INIT_FORWARD = 1, INIT_SCM = 2;      // PT.InitType ordinals (PT.java:267)
  static final int LADDER_EQUALLY_SPACED = 0, LADDER_GEOMETRIC = 1, LADDER_POLYNOMIAL = 2, LADDER_USER = 3;

  // ---- struct blangpt_config ----
  static final StructLayout CONFIG = MemoryLayout.structLayout(
      JAVA_INT.withName("nChains"), JAVA_INT.withName("nReplicas"), JAVA_LONG.withName("random"),
      JAVA_DOUBLE.withName("nPassesPerScan"), JAVA_INT.withName("usePriorSamples"), JAVA_INT.withName("reversible"),
      JAVA_INT.withName("annealSupport"), JAVA_INT.withName("treatNaNAsNegativeInfinity"), JAVA_INT.withName("device"),
      JAVA_INT.withName("rank"), JAVA_INT.withName("worldSize"), JAVA_INT.withName("ctasPerChain"),
      JAVA_LONG.withName("replicaOffset"), JAVA_INT.withName("statisticRecordedMaxChainIndex"), JAVA_INT.withName("reserved"));
  // ---- struct blangpt_array { const void* data; int64 n_elem; int32 dtype, rows, cols; (+4 padding) } ----
  static final StructLayout ARRAY = MemoryLayout.structLayout(
      ADDRESS.withName("data"), JAVA_LONG.withName("n_elem"), JAVA_INT.withName("dtype"), JAVA_INT.withName("rows"),
      JAVA_INT.withName("cols"), MemoryLayout.paddingLayout(4));
  // ---- struct blangpt_scan_outputs: six nullable pointers ----
  static final StructLayout SCAN_OUTPUTS = MemoryLayout.structLayout(
      ADDRESS.withName("energy"), ADDRESS.withName("logDensity"), ADDRESS.withName("nOutOfSupport"),
      ADDRESS.withName("swapIndicators"), ADDRESS.withName("samples"), ADDRESS.withName("intSamples"));
  // ---- struct blangpt_round_stats ----
  static final StructLayout ROUND_STATS = MemoryLayout.structLayout(
      JAVA_LONG.withName("nScansInRound"), ADDRESS.withName("swapAcceptMean"), ADDRESS.withName("energyMean"),
      ADDRESS.withName("energyVariance"), ADDRESS.withName("energyCovariance"), ADDRESS.withName("logSum"),
      ADDRESS.withName("logSumN"), ADDRESS.withName("roundTrips"), ADDRESS.withName("Lambda"),
      ADDRESS.withName("logZSteppingStone"), ADDRESS.withName("logZThermodynamic"));
  // ---- struct blangpt_diagnostics ----
  static final StructLayout DIAGNOSTICS = MemoryLayout.structLayout(
      JAVA_LONG.withName("nScansInRound"), JAVA_DOUBLE.withName("swapLowest"), JAVA_DOUBLE.withName("swapAverage"),
      JAVA_DOUBLE.withName("energyCorrelationAverage"), JAVA_DOUBLE.withName("energyCorrelationHighest"),
      JAVA_DOUBLE.withName("Lambda"), JAVA_DOUBLE.withName("inefficiency"), JAVA_DOUBLE.withName("timeToFirstRestart"),
      JAVA_DOUBLE.withName("effectiveNScans"), JAVA_LONG.withName("temperedRestarts"),
      JAVA_DOUBLE.withName("temperedRestartRate"), JAVA_DOUBLE.withName("asymptoticRoundTripCount"),
      JAVA_DOUBLE.withName("asymptoticRoundTripRate"), JAVA_DOUBLE.withName("nonAsymptoticRoundTripCount"),
      JAVA_DOUBLE.withName("nonAsymptoticRoundTripRate"), JAVA_DOUBLE.withName("logNormalizationSteppingStone"),
      JAVA_DOUBLE.withName("logNormalizationThermodynamic"),
      MemoryLayout.sequenceLayout(99, JAVA_DOUBLE).withName("cumulativeLambda"),
      MemoryLayout.sequenceLayout(99, JAVA_DOUBLE).withName("lambdaInstantaneous"));

  static long offset(StructLayout s, String field) { return s.byteOffset(PathElement.groupElement(field)); }

  // ---- functions, in the order of include/blangpt.h ----
  static final MethodHandle DEFAULT_CONFIG = fn("blangpt_default_config", FunctionDescriptor.ofVoid(ADDRESS));
  static final MethodHandle CREATE = fn("blangpt_create", i(ADDRESS, ADDRESS));
  static final MethodHandle DESTROY = fn("blangpt_destroy", FunctionDescriptor.ofVoid(ADDRESS));
  static final MethodHandle LAST_ERROR = fn("blangpt_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle SET_MODEL = fn("blangpt_set_model", i(ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT));
  static final MethodHandle N_REALS = fn("blangpt_n_reals", i(ADDRESS));
  static final MethodHandle N_INTS = fn("blangpt_n_ints", i(ADDRESS));
  static final MethodHandle N_SAMPLERS = fn("blangpt_n_samplers", i(ADDRESS));
  static final MethodHandle KERNEL_VARIANT = fn("blangpt_kernel_variant", i(ADDRESS));
  static final MethodHandle SET_LADDER = fn("blangpt_set_ladder", i(ADDRESS, ADDRESS, JAVA_LONG));
  static final MethodHandle GET_LADDER = fn("blangpt_get_ladder", i(ADDRESS, ADDRESS, JAVA_LONG));
  static final MethodHandle MAKE_LADDER = fn("blangpt_make_ladder", i(JAVA_INT, JAVA_INT, JAVA_DOUBLE, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
  static final MethodHandle SET_STATE = fn("blangpt_set_state", i(ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle GET_STATE = fn("blangpt_get_state", i(ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle SET_SCM_OPTIONS = fn("blangpt_set_scm_options", i(ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE));
  static final MethodHandle SCM_SUMMARY = fn("blangpt_scm_summary", i(ADDRESS, ADDRESS, ADDRESS, ADDRESS));
  static final MethodHandle INITIALIZE = fn("blangpt_initialize", i(ADDRESS, JAVA_INT));
  static final MethodHandle LOG_DENSITY = fn("blangpt_log_density", i(ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
  static final MethodHandle RUN_SCANS = fn("blangpt_run_scans", i(ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle GET_ROUND_STATS = fn("blangpt_get_round_stats", i(ADDRESS, ADDRESS));
  static final MethodHandle GET_DIAGNOSTICS = fn("blangpt_get_diagnostics", i(ADDRESS, JAVA_LONG, ADDRESS));
  static final MethodHandle RESET_ROUND = fn("blangpt_reset_round", i(ADDRESS));
  static final MethodHandle ADAPT = fn("blangpt_adapt", i(ADDRESS, ADDRESS));
  static final MethodHandle ADAPT_PREVIEW = fn("blangpt_adapt_preview", i(ADDRESS, ADDRESS, ADDRESS));
  static final MethodHandle ROUNDS = fn("blangpt_rounds", i(JAVA_INT, JAVA_DOUBLE, ADDRESS, ADDRESS));
  static final MethodHandle TARGET_ACCEPT_LADDER = fn("blangpt_target_accept_ladder", i(ADDRESS, JAVA_LONG, JAVA_DOUBLE, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle COUNTERS = fn("blangpt_counters", i(ADDRESS, ADDRESS));
  // multi-GPU: one handle per device inside this JVM, mailboxes connected by pointer (blangpt.h, "Multi-GPU exchange")
  static final MethodHandle PEER_MAILBOX = fn("blangpt_peer_mailbox", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle PEER_CONNECT_POINTERS = fn("blangpt_peer_connect_pointers", i(ADDRESS, ADDRESS, ADDRESS, JAVA_INT));

  /** The reference's error convention: unchecked exception with a message (Runner.xtend:164-168). */
  static void check(int status, MemorySegment handle) {
    if (status == OK) return;
    String msg;
    try {
      MemorySegment p = (MemorySegment) LAST_ERROR.invoke(handle == null ? MemorySegment.NULL : handle);
      msg = p.reinterpret(2048).getString(0);
    } catch (Throwable t) { msg = "(no message: " + t + ")"; }
    throw new RuntimeException("GpuPT [status " + status + "]: " + msg);
  }
}
