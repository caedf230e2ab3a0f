package blang.engines.gpu;

import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.List;

import blang.core.WritableIntVar;
import blang.core.WritableRealVar;

/**
 * What ModelRecogniser produces: a family of the closed catalogue (include/blangpt.h BLANGPT_FAM_*), the observed
 * data packed into contiguous host arrays, the hyper-parameters, and the model's latent variables in the order the
 * engine keeps them (blangpt_n_reals / blangpt_n_ints), so that the beta = 1 state can be written back into the Java
 * variables when a sample row is recorded (PT.recordSamples, PT.java:360-372).
 *
 * UNCOMPILED SOURCE (no JDK in the build image).
 */
public final class LoweredModel {
  public final int familyId;
  public final String familyName;
  /** data arrays in the order blangpt_set_model expects for the family; each double[] or int[] */
  public final List<Object> arrays = new ArrayList<>();
  public final double[] hyper;
  /** latent reals / ints of the model, engine order (declaration order, see blangpt.h per family) */
  public final List<WritableRealVar> reals = new ArrayList<>();
  public final List<WritableIntVar> ints = new ArrayList<>();

  LoweredModel(int familyId, String familyName, double... hyper) {
    this.familyId = familyId; this.familyName = familyName; this.hyper = hyper;
  }
  LoweredModel data(double[] a) { arrays.add(a); return this; }
  LoweredModel data(int[] a) { arrays.add(a); return this; }

  public int nArrays() { return arrays.size(); }
  public int nHyper() { return hyper.length; }
  public int nReals() { return reals.size(); }
  public int nInts() { return ints.size(); }

  /** blangpt_array[nArrays] in native memory; the library copies the data during blangpt_set_model */
  MemorySegment arraysStruct(Arena arena) {
    MemorySegment out = arena.allocate(Native.ARRAY, Math.max(1, arrays.size()));
    for (int k = 0; k < arrays.size(); k++) {
      Object a = arrays.get(k);
      MemorySegment data; long n; int dtype;
      if (a instanceof double[] d) { data = arena.allocateFrom(JAVA_DOUBLE, d); n = d.length; dtype = Native.F64; }
      else { int[] v = (int[]) a; data = arena.allocateFrom(JAVA_INT, v); n = v.length; dtype = Native.I32; }
      MemorySegment s = out.asSlice(k * Native.ARRAY.byteSize(), Native.ARRAY.byteSize());
      s.set(java.lang.foreign.ValueLayout.ADDRESS, Native.offset(Native.ARRAY, "data"), data);
      s.set(java.lang.foreign.ValueLayout.JAVA_LONG, Native.offset(Native.ARRAY, "n_elem"), n);
      s.set(JAVA_INT, Native.offset(Native.ARRAY, "dtype"), dtype);
      s.set(JAVA_INT, Native.offset(Native.ARRAY, "rows"), (int) n);
      s.set(JAVA_INT, Native.offset(Native.ARRAY, "cols"), 1);
    }
    return out;
  }
  MemorySegment hyperArray(Arena arena) { return arena.allocateFrom(JAVA_DOUBLE, hyper.length == 0 ? new double[1] : hyper); }

  /** current values of the Java variables (the prototype state: COPIES initialisation, PT.java:286-298) */
  MemorySegment currentReals(Arena arena) {
    double[] v = new double[Math.max(1, reals.size())];
    for (int k = 0; k < reals.size(); k++) v[k] = reals.get(k).doubleValue();
    return arena.allocateFrom(JAVA_DOUBLE, v);
  }
  MemorySegment currentInts(Arena arena) {
    int[] v = new int[Math.max(1, ints.size())];
    for (int k = 0; k < ints.size(); k++) v[k] = ints.get(k).intValue();
    return arena.allocateFrom(JAVA_INT, v);
  }
  /** write one recorded sample (the beta = 1 chain of scan s of a batch) back into the Java variables */
  void writeBack(ScanOutputs out, int s) {
    for (int k = 0; k < reals.size(); k++) reals.get(k).set(out.sample(s, k));
    for (int k = 0; k < ints.size(); k++) ints.get(k).set(out.intSample(s, k));
  }
}
