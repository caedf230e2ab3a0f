// families.cuh -- the closed catalogue of model families the engine runs.
//
// Each family is what the Java shim lowers a recognised Blang model to
// (INTEGRATION.md).  A family answers four questions for one chain:
//   fixed(k, x)     sum of the FIXED (prior) factors connected to sampler k's
//                   variable, evaluated with that variable set to x
//   annealed(k, x)  sum of the finite pre-annealing values of the ANNEALED
//                   factors connected to sampler k + how many of them are -inf
//                   (a group reduction over the observations)
//   commit(k, x)    write the accepted value (and refresh per-observation caches)
//   full()          preAnnealedLogLikelihood, nOutOfSupport and the fixed sum over
//                   EVERY factor (SampledModel.updateAll, R/runtime/SampledModel.java:340-364)
// Factor bodies follow R/distributions/*.bl term by term;
// what is regrouped (e.g. N identical -0.5*log(variance) terms summed as one
// product) changes results only at rounding level and is inside the 1e-9 bar.
//
// Data layout in HBM: observations are streamed with 16-byte loads
// (double2 / paired rows); per-chain parameters live in shared memory (P[]).
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace bpt {

constexpr int kMaxParams = 64;    // latent reals per chain held in shared memory
constexpr int kMaxSamplers = 65;

enum SamplerType : int { S_REAL_SLICE = 0, S_SIMPLEX = 1 };

// rows [lo, hi) of this CTA inside the chain group; lo is even
struct RowRange { int lo, hi; };
BPT_D RowRange group_rows(int n, int G, int crank) {
  int chunk = (n + G - 1) / G;
  chunk = (chunk + 1) & ~1;
  RowRange r;
  r.lo = min(n, crank * chunk);
  r.hi = min(n, r.lo + chunk);
  return r;
}

BPT_D void acc_term(double t, double& s, int& c, int* err) {
  if (t > BPT_NEG_INF) s += t;
  else { c++; if (t != t) atomicOr(err, ERR_NAN); }
}

// ---------------------------------------------------------------------------
// log-likelihood of one Bernoulli(logistic(z)) observation
// ---------------------------------------------------------------------------
// The reference evaluates the NAIVE form  log(y ? p : 1 - p),  p = 1/(1+exp(-z))
// (SURVEY.md section 10, C2).  In fp64 that form (a) returns -inf (out of support) once
// p rounds to 1.0 on a y = 0 row (z >= 36.74) and (b) loses accuracy through 1 - p for
// badly misclassified rows.  With the sign-folded margin sg = (y ? z : -z) the exact value is
//     min(sg, 0) - log1p(exp(-|sg|)).
// For sg >= -12 the naive form agrees with it to < 1e-11 absolute per row, so the hot
// path evaluates the exact form with lean fp64 kernels: the operand ranges are known, so
// there is no overflow/denormal/NaN handling, no general division and no selects.  For
// sg < -12 (or |sg| >= 700, or NaN) -- rare, and the only place where the reference's own
// rounding and its -inf convention are visible -- the naive arithmetic is executed literally.
// Coefficients live in constant memory so each one is a direct DFMA operand.
__constant__ double kExpC[12] = {   // Taylor 1/11! .. 1/0!: |r| <= ln2/2 -> relative error < 7e-15
    2.50521083854417187751e-08, 2.75573192239858906526e-07,
    2.75573192239858906526e-06, 2.48015873015873015873e-05, 1.98412698412698412698e-04,
    1.38888888888888888889e-03, 8.33333333333333333333e-03, 4.16666666666666666667e-02,
    1.66666666666666666667e-01, 0.5, 1.0, 1.0};
__constant__ double kAtanhC[8] = {  // 1/17 .. 1/3: |f| <= 0.1716 -> absolute error < 3e-16
    5.88235294117647058824e-02, 6.66666666666666666667e-02,
    7.69230769230769230769e-02, 9.09090909090909090909e-02, 1.11111111111111111111e-01,
    1.42857142857142857143e-01, 2.00000000000000000000e-01, 3.33333333333333333333e-01};
// {-log2(e), 1.5*2^52, -ln2_hi, -ln2_lo, 1 - C, 1 + C, log(C)} with C = fl(sqrt 2)
__constant__ double kLogitK[7] = {
    -1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    -0.41421356237309514547, 2.41421356237309514547, 0.34657359027997269783};

// Tables for the logistic row term (tools/gen_logit_tables.py writes them, 60-digit arithmetic, rounded once).
// exp(-a) = 2^k * 2^(j/64) * exp(r): m = round(-a * 64 log2 e) = 64 k + j, r = -a - m ln2/64, |r| <= ln2/128, so a
// degree-5 polynomial is exact to 4e-17.  log(u) for u in (1, 2]: u = c_j (1 + rr) with c_j the centre of the
// j-th of 128 mantissa intervals, rr = u * fl(1/c_j) - 1, |rr| <= 1/256, log u = -log fl(1/c_j) + log1p(rr)
// (series to rr^6, truncation < 3e-18); the table holds the log of the value really multiplied by, so the
// identity is exact up to rounding.  No division, no reciprocal; 26 fp64 instructions per row.
// {-64 log2(e), 1.5*2^52, -ln2/64 hi (21 low bits zero: m * hi exact), -ln2/64 lo}
__constant__ double kLgtK[4] = {-92.33248261689366, 6755399441055744.0, -0.01083042469326756, -2.9815858269852933e-12};
__constant__ double kLgtExpC[6] = {   // 1/5! .. 1/0!
    8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5, 1.0, 1.0};
__constant__ double kLgtLogC[5] = {-1.66666666666666666667e-01, 0.2, -0.25, 3.33333333333333333333e-01, -0.5};
__constant__ double kLgtExpTabC[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
__constant__ double2 kLgtLogTabC[128] = {
    {0.9961089494163424, 0.003898640415657309}, {0.9884169884169884, 0.01165061721997525},
    {0.9808429118773946, 0.019342962843130987}, {0.973384030418251, 0.026976587698202083},
    {0.9660377358490566, 0.03455238150665973}, {0.9588014981273408, 0.042071213920687044},
    {0.9516728624535316, 0.049533935122276676}, {0.9446494464944649, 0.05694137640013845},
    {0.9377289377289377, 0.06429435070539725}, {0.9309090909090909, 0.07159365318700882},
    {0.924187725631769, 0.078840061707776}, {0.9175627240143369, 0.08603433734180316},
    {0.9110320284697508, 0.09317722485418334}, {0.9045936395759717, 0.10026945316367517},
    {0.8982456140350877, 0.10731173578908804}, {0.89198606271777, 0.11430477128005863},
    {0.8858131487889274, 0.12124924363286965}, {0.8797250859106529, 0.12814582269193006},
    {0.8737201365187713, 0.13499516453750482}, {0.8677966101694915, 0.1417979118602574},
    {0.8619528619528619, 0.1485546943231372}, {0.8561872909698997, 0.15526612891112396},
    {0.8504983388704319, 0.16193282026931324}, {0.8448844884488449, 0.16855536102980664},
    {0.839344262295082, 0.17513433212784915}, {0.8338762214983714, 0.18167030310763463},
    {0.8284789644012945, 0.18816383241818294}, {0.8231511254019293, 0.19461546769967167},
    {0.8178913738019169, 0.2010257460605908}, {0.8126984126984127, 0.2073951943460706},
    {0.807570977917981, 0.21372432939771818}, {0.8025078369905956, 0.22001365830528213},
    {0.7975077881619937, 0.2262636786504534}, {0.7925696594427245, 0.232474878743094},
    {0.7876923076923077, 0.238647737850175}, {0.7828746177370031, 0.24478272641769092},
    {0.7781155015197568, 0.25088030628580943}, {0.7734138972809668, 0.2569409308975004},
    {0.7687687687687688, 0.26296504550088134}, {0.764179104477612, 0.26895308734550394},
    {0.7596439169139466, 0.2749054858727992}, {0.7551622418879056, 0.2808226629008878},
    {0.750733137829912, 0.2867050328039543}, {0.7463556851311953, 0.29255300268637746},
    {0.7420289855072464, 0.2983669725517973}, {0.7377521613832853, 0.3041473354672968},
    {0.7335243553008596, 0.3098944777228647}, {0.7293447293447294, 0.3156087789863033},
    {0.7252124645892352, 0.32129061245373425}, {0.7211267605633803, 0.3269403449958533},
    {0.7170868347338936, 0.3325583373000766}, {0.713091922005571, 0.3381449440087164},
    {0.7091412742382271, 0.34370051385331846}, {0.7052341597796143, 0.3492253897852883},
    {0.7013698630136986, 0.354719909102929}, {0.6975476839237057, 0.3601844035750078},
    {0.6937669376693767, 0.3656191995609647}, {0.6900269541778976, 0.37102461812787263},
    {0.6863270777479893, 0.376400975164253}, {0.6826666666666666, 0.3817485814908484},
    {0.6790450928381963, 0.3870677429684483}, {0.6754617414248021, 0.3923587606028639},
    {0.6719160104986877, 0.3976219306471385}, {0.6684073107049608, 0.4028575447010835},
    {0.6649350649350649, 0.4080658898082217}, {0.661498708010336, 0.41324724855021927},
    {0.6580976863753213, 0.41840189913888387}, {0.6547314578005116, 0.4235301155058032},
    {0.6513994910941476, 0.42863216738969867}, {0.6481012658227848, 0.4337083204215594},
    {0.6448362720403022, 0.43875883620762796}, {0.6416040100250626, 0.44378397241030104},
    {0.6384039900249376, 0.4487839828270067}, {0.6352357320099256, 0.4537591174671205},
    {0.6320987654320988, 0.4587096226269767}, {0.628992628992629, 0.46363574096303256},
    {0.6259168704156479, 0.46853771156323926}, {0.6228710462287105, 0.4734157700166721},
    {0.6198547215496368, 0.47827014848147026}, {0.6168674698795181, 0.48310107575113576},
    {0.6139088729016786, 0.48790877731923904}, {0.6109785202863962, 0.4926934754425752},
    {0.6080760095011877, 0.4974553892028189}, {0.6052009456264775, 0.5021947345667155},
    {0.6023529411764705, 0.5069117244448544}, {0.5995316159250585, 0.5116065687490621},
    {0.5967365967365967, 0.5162794744484545}, {0.5939675174013921, 0.5209306456241853},
    {0.5912240184757506, 0.5255602835229274}, {0.5885057471264368, 0.5301685866091216},
    {0.585812356979405, 0.5347557506160276}, {0.5831435079726651, 0.5393219685956089},
    {0.5804988662131519, 0.5438674309672835}, {0.5778781038374717, 0.5483923255655733},
    {0.5752808988764045, 0.5528968376866776}, {0.5727069351230425, 0.5573811501340064},
    {0.5701559020044543, 0.5618454432626918}, {0.5676274944567627, 0.5662898950231159},
    {0.565121412803532, 0.5707146810034716}, {0.5626373626373626, 0.575119974471388},
    {0.5601750547045952, 0.5795059464146423}, {0.5577342047930284, 0.5838727655809826},
    {0.5553145336225597, 0.588220598517086}, {0.5529157667386609, 0.5925496096066716},
    {0.5505376344086022, 0.5968599611077938}, {0.5481798715203426, 0.6011518131893347},
    {0.5458422174840085, 0.6054253239667169}, {0.5435244161358811, 0.6096806495368553},
    {0.5412262156448203, 0.6139179440123704}, {0.5389473684210526, 0.6181373595550788},
    {0.5366876310272537, 0.6223390464087787}, {0.534446764091858, 0.6265231529313529},
    {0.5322245322245323, 0.6306898256261987}, {0.5300207039337475, 0.6348392091730102},
    {0.5278350515463918, 0.6389714464579207}, {0.5256673511293635, 0.6430866786030273},
    {0.523517382413088, 0.6471850449953095}, {0.5213849287169042, 0.6512666833149582},
    {0.5192697768762677, 0.6553317295631277}, {0.5171717171717172, 0.6593803180891278},
    {0.5150905432595574, 0.6634125816170662}, {0.5130260521042084, 0.6674286512719563},
    {0.5109780439121756, 0.6714286566053024}, {0.5089463220675944, 0.6754127256201768},
    {0.5069306930693069, 0.6793809847957973}, {0.504930966469428, 0.6833335591116206},
    {0.5029469548133595, 0.6872705720709603}, {0.5009784735812133, 0.691192145724142}};

// the tables live in shared memory (lanes index them with different j); one copy per CTA, filled once
struct LogitTables { double2 lg[128]; double ex[64]; };
__shared__ LogitTables s_logit_tab;
BPT_D void logit_table_init() {   // call before the first __syncthreads() of the kernel
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_logit_tab.lg[i] = kLgtLogTabC[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_logit_tab.ex[i] = kLgtExpTabC[i];
}

BPT_D uint32_t logit_table_address() {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(&s_logit_tab);
  asm volatile("" : "+r"(a));   // opaque: keeps the address in a register instead of re-deriving it per use
  return a;
}
// W independent lanes evaluated together: min(z,0) - log1p(exp(-|z|)) for -12 <= z, |z| < 700.
// Straight-line code; the W chains interleave (ILP) and share every coefficient.  Inputs outside the range
// produce garbage but never fault (callers may evaluate rows they will discard).
template <int W>
BPT_D void logit_fast(const double (&z)[W], double& s, uint32_t tab) {
  double kd[W], r[W], p[W], te[W]; int m[W];
#pragma unroll
  for (int w = 0; w < W; w++) {
    const double a = fabs(z[w]);
    kd[w] = fma(a, kLgtK[0], kLgtK[1]);     // magic add: low word = round(-a * 64 log2 e)
    m[w] = __double2loint(kd[w]);
    kd[w] -= kLgtK[1];
    asm("ld.shared.f64 %0, [%1];" : "=d"(te[w]) : "r"(tab + 2048u + ((uint32_t)(m[w] & 63) << 3)));
    r[w] = fma(kd[w], kLgtK[2], -a);
    r[w] = fma(kd[w], kLgtK[3], r[w]);
    p[w] = kLgtExpC[0];
  }
#pragma unroll
  for (int i = 1; i < 6; i++)
#pragma unroll
    for (int w = 0; w < W; w++) p[w] = fma(p[w], r[w], kLgtExpC[i]);
  double q[W], rr[W], lc[W];
#pragma unroll
  for (int w = 0; w < W; w++) {
    // t = exp(-|z|) = 2^k * (2^(j/64) * p) in (0, 1];  u = 1 + t in (1, 2]
    const double tp = te[w] * p[w];
    const double t = __hiloint2double(__double2hiint(tp) + ((m[w] >> 6) << 20), __double2loint(tp));
    const double u = 1.0 + t;
    const unsigned j = min(127u, (unsigned)(__double2hiint(u) - 0x3ff00000) >> 13);
    double2 cj;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cj.x), "=d"(cj.y) : "r"(tab + (j << 4)));
    rr[w] = fma(u, cj.x, -1.0);
    lc[w] = cj.y;
    q[w] = kLgtLogC[0];
  }
#pragma unroll
  for (int i = 1; i < 5; i++)
#pragma unroll
    for (int w = 0; w < W; w++) q[w] = fma(q[w], rr[w], kLgtLogC[i]);
#pragma unroll
  for (int w = 0; w < W; w++) {
    const double l = fma(rr[w] * rr[w], q[w], rr[w]) + lc[w];     // log1p(exp(-|z|))
    s += fma(0.5, z[w] - fabs(z[w]), -l);                           // min(z,0) - log1p(exp(-|z|))
  }
}
BPT_D bool logit_is_fast(double zs) { return zs >= -12.0 && fabs(zs) < 700.0; }

// the literal reference arithmetic (rounding + -inf convention); NaN lands here too
BPT_D void logit_slow(double zs, const uint8_t* yp, double& s, int& c, int* err) {
  const int y = *yp;
  const double z = y ? zs : -zs;
  const double pr = 1.0 / (1.0 + exp(-z));
  const double pick = y ? pr : 1.0 - pr;
  acc_term(log(pick), s, c, err);
}
// zs = sign-folded margin (y ? z : -z); *yp is only read on the rare literal path
BPT_D void logit_term(double zs, const uint8_t* yp, double& s, int& c, int* err) {
  if (logit_is_fast(zs)) { const double z[1] = {zs}; logit_fast<1>(z, s, logit_table_address()); }
  else logit_slow(zs, yp, s, c, err);
}
// Two rows.  The lean arithmetic runs unconditionally (straight-line, no divergence bookkeeping on the hot
// path); its result is kept when both rows are in its range, which one warp vote decides for all lanes that
// are here together.  Otherwise each row is redone on its own path.
template <bool FULL_WARP = false>
BPT_D void logit_pair(double za, double zb, const uint8_t* yp, double& s, int& c, int* err, uint32_t tab) {
  const bool ok = logit_is_fast(za) && logit_is_fast(zb);
  const double z[2] = {za, zb};
  double t = s;
  logit_fast<2>(z, t, tab);
  // FULL_WARP: the caller guarantees that all 32 lanes are here together
  if (__all_sync(FULL_WARP ? 0xffffffffu : __activemask(), ok)) { s = t; return; }
  if (ok) s = t;
  else { logit_term(za, yp, s, c, err); logit_term(zb, yp + 1, s, c, err); }
}

// ---------------------------------------------------------------------------
// lean fp64 exp / log for the mixture kernel (operand ranges known, see MixtureFam::obs_fast)
// ---------------------------------------------------------------------------
BPT_D double exp_lean(double x) {   // exp(x) for -700 <= x <= 700 (caller clamps)
  double kd = fma(x, -kLogitK[0], kLogitK[1]);
  const int k = __double2loint(kd);
  kd -= kLogitK[1];
  double r = fma(kd, kLogitK[2], x);
  r = fma(kd, kLogitK[3], r);
  double p = kExpC[0];
#pragma unroll
  for (int i = 1; i < 12; i++) p = fma(p, r, kExpC[i]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
BPT_D double log_lean(double a) {   // log(a) for normal positive a
  const int hi = __double2hiint(a);
  int mh = (hi & 0x000fffff) | 0x3ff00000;          // mantissa in [1, 2)
  const int adj = mh >= 0x3ff6a09f ? 1 : 0;          // above sqrt(2): halve
  mh -= adj << 20;
  const int e = (hi >> 20) - 1023 + adj;
  const double m = __hiloint2double(mh, __double2loint(a));   // [0.7071, 1.4142)
  const double num = m - 1.0, den = m + 1.0;
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
  double t = fma(-den, rc, 1.0); rc = fma(rc, t, rc);
  t = fma(-den, rc, 1.0); rc = fma(rc, t, rc);
  const double f = num * rc, f2 = f * f;
  double q = kAtanhC[0];
#pragma unroll
  for (int i = 1; i < 8; i++) q = fma(q, f2, kAtanhC[i]);
  const double s1 = f + f;
  return fma((double)e, 6.93147180559945309417e-01, fma(s1 * f2, q, s1));
}

// lean branch-free lnGamma for the beta-binomial kernel: shift by 8 and use Stirling's series at z = x + 8 >= 8
//   lnGamma(x) = (z - 1/2) log z - z + log(2 pi)/2 + sum_j B_2j / (2j (2j-1) z^(2j-1)) - log( x (x+1) ... (x+7) )
// (series truncated after z^-13: error < 1e-15).  All lanes execute the same instructions whatever their
// argument, unlike the library lgamma whose range branches serialise inside a warp.  Valid for 1e-30 < x < 1e15;
// outside, the library function is used.
BPT_D double lgamma_lean(double x) {
  if (!(x > 1e-30 && x < 1e15)) return lgamma(x);
  double p = x;
#pragma unroll
  for (int j = 1; j < 8; j++) p *= x + (double)j;
  const double z = x + 8.0;
  double rz;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rz) : "d"(z));
  double t = fma(-z, rz, 1.0); rz = fma(rz, t, rz);
  t = fma(-z, rz, 1.0); rz = fma(rz, t, rz);
  const double r2 = rz * rz;
  double s = 6.41025641025641025641e-03;            //  1/156
  s = fma(s, r2, -1.91752691752691752692e-03);      // -691/360360
  s = fma(s, r2, 8.41750841750841750842e-04);       //  1/1188
  s = fma(s, r2, -5.95238095238095238095e-04);      // -1/1680
  s = fma(s, r2, 7.93650793650793650794e-04);       //  1/1260
  s = fma(s, r2, -2.77777777777777777778e-03);      // -1/360
  s = fma(s, r2, 8.33333333333333333333e-02);       //  1/12
  const double body = fma(z - 0.5, log_lean(z), -z) + 9.18938533204672741780e-01;
  return fma(s, rz, body) - log_lean(p);
}

// Normal.bl:21-24 with constant mean/variance (a prior on x)
BPT_D double normal_f2(double mean, double variance, double x) {
  if (variance <= 0.0) return BPT_NEG_INF;
  const double d = mean - x;
  return -0.5 * (d * d) / variance;
}
// Gamma.bl:14-23, the two factors that read the realization
BPT_D double gamma_ab(double shape, double rate, double x) {
  if (shape <= 0.0 || rate <= 0.0) return BPT_NEG_INF;
  if (x <= 0.0) return BPT_NEG_INF;
  return (shape - 1.0) * log(x * rate) + (-x * rate);
}
// Gamma.bl:24-31, the two constant factors
BPT_D double gamma_cd(double shape, double rate) {
  if (shape <= 0.0 || rate <= 0.0) return BPT_NEG_INF;
  return -lgamma(shape) + log(rate);
}

// ===========================================================================
// C1  Normal with unknown mean and variance.  P = [mu, sigma2]
// ===========================================================================
struct NormalFam {
  struct Data {
    const double* y; int n;
    double m0, v0, rate;
  };
  static constexpr int kThreads = 128;
  static constexpr bool kWarpCapable = true;   // may run one chain per warp in the fused replica kernel
  Data d; double* P; RowRange rr; int* err;
  int t0, tn;   // this thread's index in the chain group and the group size

  BPT_D static int n_reals(const Data&) { return 2; }
  BPT_D static int n_samplers(const Data&) { return 2; }
  BPT_D void init(const Data& d_, int, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; rr = group_rows(d.n, G, crank); err = err_; t0 = threadIdx.x; tn = blockDim.x;
  }
  BPT_D void set_lanes(int lane, int width) { t0 = lane; tn = width; }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  BPT_D double fixed(int k, double x, int) const {
    return k == 0 ? normal_f2(d.m0, d.v0, x) : gamma_ab(1.0, d.rate, x);   // Exponential.bl:11 -> Gamma(1.0, rate)
  }
  BPT_D double sum_sq(double mu) const {
    double s = 0.0;
    for (int i = rr.lo + 2 * t0; i < rr.hi; i += 2 * tn) {
      if (i + 1 < rr.hi) {
        const double2 v = *reinterpret_cast<const double2*>(d.y + i);
        const double a = mu - v.x, b = mu - v.y;
        s = fma(a, a, s); s = fma(b, b, s);
      } else { const double a = mu - d.y[i]; s = fma(a, a, s); }
    }
    return s;
  }
  // mu: the n F2 factors; sigma2: the n F1 and the n F2 factors (Normal.bl:17-24)
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double mu = k == 0 ? x : P[0], s2 = k == 1 ? x : P[1];
    if (s2 <= 0.0) { s = 0.0; c = (k == 0 ? 1 : 2) * d.n; return; }
    double q = sum_sq(mu); int z = 0;
    red.reduce(q, z);
    s = -0.5 * q / s2;
    if (k == 1) s += d.n * (-0.5 * log(s2));
    c = 0;
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    const double mu = P[0], s2 = P[1];
    loglik = d.n * (-0.5 * BPT_LOG_2PI);
    if (s2 <= 0.0) noos = 2 * d.n;
    else {
      double q = sum_sq(mu); int z = 0;
      red.reduce(q, z);
      loglik += d.n * (-0.5 * log(s2)) + (-0.5 * q / s2);
      noos = 0;
    }
    logprior = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0)) + normal_f2(d.m0, d.v0, mu) +
               gamma_ab(1.0, d.rate, s2) + gamma_cd(1.0, d.rate);
  }
  BPT_D void forward(Rng& rng) {   // declaration order: mu then sigma2
    const double mu = gen_normal(rng, d.m0, d.v0);
    const double s2 = gen_exponential(rng, d.rate);
    P[0] = mu; P[1] = s2;
  }
};

// ===========================================================================
// C2  Bayesian logistic regression.  P = w[0..p)
//     per-chain cache eta_i = s_i * (x_i . w) in HBM, s_i = +1 if y_i = 1 else -1 (the rows of
//     the device copy of X are sign-folded once at set_model, which is exact); a move of w_k
//     needs only column k of X:  eta_i(x) = eta_i + (x - w_k) * s_i X[i][k]
// ===========================================================================
template <int THREADS>
struct LogisticFamT {
  struct Data {
    const double* xt;      // [p][npad] column-major, sign-folded copy of the design matrix
    const uint8_t* y;      // [npad]
    double* eta;           // [localSlots][npad]
    int n, npad, p;
    double v0;
  };
  static constexpr int kThreads = THREADS;
  static constexpr bool kWarpCapable = false;
  Data d; double* P; double* eta; RowRange rr; int* err;

  BPT_D static int n_reals(const Data& d) { return d.p; }
  BPT_D static int n_samplers(const Data& d) { return d.p; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; eta = d.eta + (size_t)local_slot * d.npad; rr = group_rows(d.n, G, crank); err = err_;
  }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() { logit_table_init(); }
  BPT_D double fixed(int, double x, int) const { return normal_f2(0.0, d.v0, x); }

  // Bernoulli.bl:11-14 -> Categorical.bl:12-15 with p = logistic(eta).  See logit_term() above.
  BPT_D void row_term(double zs, const uint8_t* yp, double& s, int& c) const { logit_term(zs, yp, s, c, err); }

  // one pass over this CTA's rows: eta and column k of X streamed as 16-byte pairs, the next pair
  // prefetched before the current pair's ~90 fp64 instructions; both rows of a pair run interleaved
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double delta = x - P[k];
    double acc = 0.0; int cnt = 0;
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    const double2* pe = reinterpret_cast<const double2*>(eta + rr.lo) + threadIdx.x;
    const double2* px = reinterpret_cast<const double2*>(d.xt + (size_t)k * d.npad + rr.lo) + threadIdx.x;
    const uint8_t* py = d.y + rr.lo + 2 * threadIdx.x;
    const uint32_t tab = logit_table_address();
    // `whole` counts the pairs from this warp's first lane on: while it is >= 32 all lanes of the warp have a
    // pair, so the loop is warp-uniform (full-mask vote inside logit_pair); the ragged rest is one more step.
    int whole = n_pairs - (int)(threadIdx.x & ~31u);
    // two register sets alternate between "being evaluated" and "being loaded" (no copies between them)
    double2 ea = make_double2(0.0, 0.0), xa = ea, eb = ea, xb = ea;
    if (whole >= 32) { ea = *pe; xa = *px; }
    while (whole >= 32) {
      whole -= THREADS;
      if (whole >= 32) { eb = pe[THREADS]; xb = px[THREADS]; }
      logit_pair<true>(fma(delta, xa.x, ea.x), fma(delta, xa.y, ea.y), py, acc, cnt, err, tab);
      pe += THREADS; px += THREADS; py += 2 * THREADS;
      if (whole < 32) break;
      whole -= THREADS;
      if (whole >= 32) { ea = pe[THREADS]; xa = px[THREADS]; }
      logit_pair<true>(fma(delta, xb.x, eb.x), fma(delta, xb.y, eb.y), py, acc, cnt, err, tab);
      pe += THREADS; px += THREADS; py += 2 * THREADS;
    }
    if (whole - (int)(threadIdx.x & 31u) > 0) {      // last, partly filled warp step
      const double2 e = *pe, xv = *px;
      logit_pair<false>(fma(delta, xv.x, e.x), fma(delta, xv.y, e.y), py, acc, cnt, err, tab);
    }
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) {   // odd tail row of this CTA's range
      const int i = rr.hi - 1;
      row_term(fma(delta, d.xt[(size_t)k * d.npad + i], eta[i]), d.y + i, acc, cnt);
    }
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int) {
    const double delta = x - P[k];
    const double* col = d.xt + (size_t)k * d.npad;
    if (delta != 0.0)
      for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * THREADS) {
        if (i + 1 < rr.hi) {
          double2 e = *reinterpret_cast<const double2*>(eta + i);
          const double2 xv = *reinterpret_cast<const double2*>(col + i);
          e.x = fma(delta, xv.x, e.x); e.y = fma(delta, xv.y, e.y);
          *reinterpret_cast<double2*>(eta + i) = e;
        } else eta[i] = fma(delta, col[i], eta[i]);
      }
    __syncthreads();   // every thread has read P[k] for its delta before anyone overwrites it
    P[k] = x;
  }
  // rebuild eta from scratch: sum_j w_j * X[i][j], j ascending from 0.0 (F/SpikedGLM.bl:29-33)
  BPT_D void refresh() {
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * THREADS) {
      const bool two = i + 1 < rr.hi;
      double a = 0.0, b = 0.0;
      for (int j = 0; j < d.p; j++) {
        const double w = P[j];
        const double* col = d.xt + (size_t)j * d.npad;
        if (two) { const double2 xv = *reinterpret_cast<const double2*>(col + i); a = fma(w, xv.x, a); b = fma(w, xv.y, b); }
        else a = fma(w, col[i], a);
      }
      eta[i] = a; if (two) eta[i + 1] = b;
    }
  }
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) {
    refresh();
    double acc = 0.0; int cnt = 0;
    for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * THREADS) {
      row_term(eta[i], d.y + i, acc, cnt);
      if (i + 1 < rr.hi) row_term(eta[i + 1], d.y + i + 1, acc, cnt);
    }
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    double lp = 0.0;
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int j = 0; j < d.p; j++) lp += f01 + normal_f2(0.0, d.v0, P[j]);
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {
    for (int j = 0; j < d.p; j++) { const double w = gen_normal(rng, 0.0, d.v0); P[j] = w; }
  }
};
using LogisticFam = LogisticFamT<512>;

// ===========================================================================
// C3  K-component Gaussian mixture, indicators marginalised.
//     P = [mu[K], var[K], pi[K]];  samplers: k<K mu_k, K<=k<2K var_k, 2K simplex
//     per observation: log( sum_k pi_k * exp(logN_k(y_i)) )   (naive form, F/Multimodal.bl:12-14)
// ===========================================================================
template <int KT>
struct MixtureFam {
  struct Data {
    const double* y; int n; int K;
    double m0, v0, gshape, grate, conc;
    // per-chain scratch in HBM, [localSlots][npad] each: within one sampler execution only one component
    // (or the two simplex entries) changes, so the first evaluation stores the sum of the OTHER components'
    // terms per observation and the remaining ~8 evaluations need one exp (or none) + one log per row
    double* rest; double* e1; double* e2; int npad;
  };
  static constexpr int kThreads = 512;
  static constexpr bool kWarpCapable = false;
  Data d; double* P; double* Cst; RowRange rr; int* err;   // Cst: [4][KT] per-evaluation component constants
  double *rest, *e1, *e2;
  bool first;   // next annealed() call is the first of a sampler execution

  BPT_D static int n_reals(const Data& d) { return 3 * d.K; }
  BPT_D static int n_samplers(const Data& d) { return 2 * d.K + 1; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; Cst = P_ + kMaxParams; rr = group_rows(d.n, G, crank); err = err_;
    rest = d.rest + (size_t)local_slot * d.npad; e1 = d.e1 + (size_t)local_slot * d.npad;
    e2 = d.e2 + (size_t)local_slot * d.npad; first = true;
  }
  BPT_D void begin_execution() { first = true; }
  BPT_D static void tables_init() {}
  BPT_D int sampler_type(int k) const { return k == 2 * d.K ? S_SIMPLEX : S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return d.K; }
  // candidate-substituted parameter of component c
  BPT_D void substituted(int k, double x, int aux, int c, double& mu, double& var, double& pi) const {
    const int K = d.K;
    mu = P[c]; var = P[K + c]; pi = P[2 * K + c];
    if (k < K) { if (c == k) mu = x; }
    else if (k < 2 * K) { if (c == k - K) var = x; }
    else {   // SimplexWritableVariable.set :26-30
      const int nx = aux == K - 1 ? 0 : aux + 1;
      if (c == aux) pi = x;
      else if (c == nx) pi = fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x);
    }
  }
  BPT_D double fixed(int k, double x, int aux) const {
    const int K = d.K;
    if (k < K) return normal_f2(d.m0, d.v0, x);
    if (k < 2 * K) return gamma_ab(d.gshape, d.grate, x);
    double sum = 0.0;   // Dirichlet.bl:13-22
    for (int c = 0; c < K; c++) {
      double mu, var, pi; substituted(k, x, aux, c, mu, var, pi);
      if (d.conc < 0.0) return BPT_NEG_INF;
      sum += (d.conc - 1.0) * log(pi);
    }
    return sum;
  }
  BPT_D void stage_constants(int k, double x, int aux) {
    __syncthreads();
    if (threadIdx.x < KT) {
      const int c = threadIdx.x;
      double mu = 0.0, var = 1.0, pi = 0.0;      // padding components contribute exactly 0
      if (c < d.K) substituted(k, x, aux, c, mu, var, pi);
      Cst[0 * KT + c] = mu;
      Cst[1 * KT + c] = var <= 0.0 ? BPT_NEG_INF : -0.5 * BPT_LOG_2PI + (-0.5 * log(var));
      Cst[2 * KT + c] = var <= 0.0 ? 0.0 : -0.5 / var;
      Cst[3 * KT + c] = pi;
    }
    __syncthreads();
  }
  // literal form: log( sum_k pi_k * exp(logN_k(y)) ) with library exp/log (handles -inf, 0, NaN)
  BPT_D void obs_slow(double yv, const double* mu, const double* a, const double* b, const double* pi, double& s,
                      int& c) const {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < KT; q++) { const double dd = mu[q] - yv; acc += pi[q] * exp(fma(b[q], dd * dd, a[q])); }
    acc_term(log(acc), s, c, err);
  }
  // hot path: the KT exponentials run as KT interleaved chains; arguments are clamped to >= -700
  // (a clamped term is < 1e-304, negligible next to the accepted range acc > 1e-250); anything
  // outside that range -- every component underflowing, invalid variances, NaN -- is redone literally
  BPT_D void obs_term(double yv, const double* mu, const double* a, const double* b, const double* pi, double& s,
                      int& c) const {
    double acc = 0.0;
#pragma unroll
    for (int q0 = 0; q0 < KT; q0 += 4) {     // four interleaved exp chains at a time
      double arg[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const double dd = mu[q0 + q] - yv; arg[q] = fmax(fma(b[q0 + q], dd * dd, a[q0 + q]), -700.0); }
#pragma unroll
      for (int q = 0; q < 4; q++) acc = fma(pi[q0 + q], exp_lean(arg[q]), acc);
    }
    if (acc > 1e-250 && acc < 1e250) s += log_lean(acc);
    else obs_slow(yv, mu, a, b, pi, s, c);
  }
  // not inlined: the loop gets its own register allocation instead of competing with the sampler state
  __device__ __noinline__ void sum_obs(double& s, int& c) const {
    double mu[KT], a[KT], b[KT], pi[KT];
#pragma unroll
    for (int q = 0; q < KT; q++) { mu[q] = Cst[q]; a[q] = Cst[KT + q]; b[q] = Cst[2 * KT + q]; pi[q] = Cst[3 * KT + q]; }
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    const double2* py = reinterpret_cast<const double2*>(d.y + rr.lo) + threadIdx.x;
    int left = n_pairs - (int)threadIdx.x;
    double2 vn = make_double2(0.0, 0.0);
    if (left > 0) vn = *py;
    while (left > 0) {
      const double2 v = vn;
      left -= blockDim.x; py += blockDim.x;
      if (left > 0) vn = *py;
      obs_term(v.x, mu, a, b, pi, s, c);
      obs_term(v.y, mu, a, b, pi, s, c);
    }
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) obs_term(d.y[rr.hi - 1], mu, a, b, pi, s, c);
  }
  // First evaluation of a sampler execution (always at the current state): every component is evaluated, and
  // per observation the sum R of the components the sampler does NOT move is stored (plus, for the simplex
  // sampler, the two moved components' unweighted densities).  j0 / j1 = moved components (j1 = -1 if one).
  __device__ __noinline__ void pass_first(int j0, int j1, double& s, int& c) const {
    double mu[KT], a[KT], b[KT], pi[KT];
#pragma unroll
    for (int q = 0; q < KT; q++) { mu[q] = Cst[q]; a[q] = Cst[KT + q]; b[q] = Cst[2 * KT + q]; pi[q] = Cst[3 * KT + q]; }
    auto one = [&](double yv, double& r_out, double& e1_out, double& e2_out) {
      double R = 0.0, M = 0.0, E1 = 0.0, E2 = 0.0;
#pragma unroll
      for (int q0 = 0; q0 < KT; q0 += 4) {
        double arg[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const double dd = mu[q0 + q] - yv; arg[q] = fmax(fma(b[q0 + q], dd * dd, a[q0 + q]), -700.0); }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double e = exp_lean(arg[q]);
          const bool moved = (q0 + q == j0) || (q0 + q == j1);
          if (q0 + q == j0) E1 = e;
          if (q0 + q == j1) E2 = e;
          if (moved) M = fma(pi[q0 + q], e, M); else R = fma(pi[q0 + q], e, R);
        }
      }
      const double S = R + M;
      if (S > 1e-250 && S < 1e250) { s += log_lean(S); r_out = R; }
      else { obs_slow(yv, mu, a, b, pi, s, c); r_out = CUDART_NAN; }   // NaN marks a row the cached passes redo literally
      e1_out = E1; e2_out = E2;
    };
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
      const int i = rr.lo + 2 * p;
      const double2 v = *reinterpret_cast<const double2*>(d.y + i);
      double2 r, x1, x2;
      one(v.x, r.x, x1.x, x2.x); one(v.y, r.y, x1.y, x2.y);
      *reinterpret_cast<double2*>(rest + i) = r;
      if (j1 >= 0) { *reinterpret_cast<double2*>(e1 + i) = x1; *reinterpret_cast<double2*>(e2 + i) = x2; }
    }
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) {
      const int i = rr.hi - 1;
      double r, x1, x2;
      one(d.y[i], r, x1, x2);
      rest[i] = r; if (j1 >= 0) { e1[i] = x1; e2[i] = x2; }
    }
  }
  // Remaining evaluations of the execution: S_i = R_i + (moved terms at the candidate).  One exp + one log per
  // row for a mean/variance move (SIMPLEX = false), no exp at all for a simplex move; 16 (or 32) bytes per row
  // streamed from HBM with no reuse inside an evaluation, so kB pairs per thread are in flight and the next kB
  // already requested while the current ones are evaluated (the pass is latency-, not bandwidth-bound).
  template <bool SIMPLEX, int kB>
  __device__ __noinline__ void pass_cached(int j0, int j1, double& s, int& c) const {
    const double mu0 = Cst[j0], a0 = Cst[KT + j0], b0 = Cst[2 * KT + j0], pi0 = Cst[3 * KT + j0];
    const double pi1 = SIMPLEX ? Cst[3 * KT + j1] : 0.0;
    auto slow = [&](double yv) {
      double mu[KT], a[KT], b[KT], pi[KT];
#pragma unroll
      for (int q = 0; q < KT; q++) { mu[q] = Cst[q]; a[q] = Cst[KT + q]; b[q] = Cst[2 * KT + q]; pi[q] = Cst[3 * KT + q]; }
      obs_slow(yv, mu, a, b, pi, s, c);
    };
    auto one = [&](double yv, double R, double E1, double E2) {
      double S;
      if (!SIMPLEX) { const double dd = mu0 - yv; S = fma(pi0, exp_lean(fmax(fma(b0, dd * dd, a0), -700.0)), R); }
      else S = fma(pi0, E1, fma(pi1, E2, R));
      if (S > 1e-250 && S < 1e250) s += log_lean(S);
      else slow(yv);
    };
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    const int stride = blockDim.x;
    int p = threadIdx.x;
    double2 vn[kB], rn[kB], an[SIMPLEX ? kB : 1], bn[SIMPLEX ? kB : 1];
    auto load = [&](int pp) {
#pragma unroll
      for (int u = 0; u < kB; u++) {
        const int q = pp + u * stride;
        if (q < n_pairs) {
          const int i = rr.lo + 2 * q;
          vn[u] = *reinterpret_cast<const double2*>(d.y + i); rn[u] = __ldcs(reinterpret_cast<const double2*>(rest + i));
          if (SIMPLEX) { an[u] = __ldcs(reinterpret_cast<const double2*>(e1 + i)); bn[u] = __ldcs(reinterpret_cast<const double2*>(e2 + i)); }
        }
      }
    };
#pragma unroll
    for (int u = 0; u < kB; u++) { vn[u] = make_double2(0.0, 0.0); rn[u] = vn[u]; }
    an[0] = make_double2(0.0, 0.0); bn[0] = an[0];
    if (p < n_pairs) load(p);
    while (p < n_pairs) {
      double2 v[kB], r[kB], x1[SIMPLEX ? kB : 1], x2[SIMPLEX ? kB : 1];
#pragma unroll
      for (int u = 0; u < kB; u++) { v[u] = vn[u]; r[u] = rn[u]; if (SIMPLEX) { x1[u] = an[u]; x2[u] = bn[u]; } }
      if (!SIMPLEX) { x1[0] = an[0]; x2[0] = bn[0]; }
      const int cur = p;
      p += kB * stride;
      if (p < n_pairs) load(p);
#pragma unroll
      for (int u = 0; u < kB; u++)
        if (cur + u * stride < n_pairs) {
          const int w = SIMPLEX ? u : 0;
          one(v[u].x, r[u].x, x1[w].x, x2[w].x); one(v[u].y, r[u].y, x1[w].y, x2[w].y);
        }
    }
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) {
      const int i = rr.hi - 1;
      one(d.y[i], rest[i], SIMPLEX ? e1[i] : 0.0, SIMPLEX ? e2[i] : 0.0);
    }
  }
  BPT_D void annealed(int k, double x, int aux, GroupReducer& red, double& s, int& c) {
    const int K = d.K;
    const int j0 = k < K ? k : (k < 2 * K ? k - K : aux);
    const int j1 = k < 2 * K ? -1 : (aux == K - 1 ? 0 : aux + 1);
    stage_constants(k, x, aux);
    double acc = 0.0; int cnt = 0;
    if (first) { pass_first(j0, j1, acc, cnt); first = false; }
    else if (j1 < 0) pass_cached<false, 4>(j0, j1, acc, cnt);
    else pass_cached<true, 2>(j0, j1, acc, cnt);
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int aux) {
    const int K = d.K;
    if (k < 2 * K) { P[k] = x; return; }
    const int nx = aux == K - 1 ? 0 : aux + 1;
    const double comp = fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x);
    __syncthreads();
    P[2 * K + aux] = x; P[2 * K + nx] = comp;
    __syncthreads();
  }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) {
    const int K = d.K;
    stage_constants(0, P[0], 0);
    double acc = 0.0; int cnt = 0;
    sum_obs(acc, cnt);
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    double lp = 0.0;
    // F/MixtureModel.bl:14-20 order: Dirichlet, then per component Normal, Gamma
    for (int c = 0; c < K; c++) lp += (d.conc - 1.0) * log(P[2 * K + c]);
    lp += d.conc < 0.0 ? BPT_NEG_INF : (-K * lgamma(d.conc) + lgamma(K * d.conc));
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int c = 0; c < K; c++) {
      lp += f01 + normal_f2(d.m0, d.v0, P[c]);
      lp += gamma_ab(d.gshape, d.grate, P[K + c]) + gamma_cd(d.gshape, d.grate);
    }
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {
    const int K = d.K;
    double g[KT]; double sum = 0.0;
    // Generators.dirichletInPlace :233-272
    if (d.conc < 1.0) {
      double lse = BPT_NEG_INF;
      for (int c = 0; c < K; c++) {
        const double gv = gen_gamma(rng, d.conc + 1.0, 1.0);
        const double logu = log(rng.next_double());
        g[c] = log(gv) + logu / d.conc;
        lse = log_add(lse, g[c]);
      }
      for (int c = 0; c < K; c++) g[c] = exp(g[c] - lse);
    } else {
      for (int c = 0; c < K; c++) { g[c] = gen_gamma(rng, d.conc, 1e7); sum += g[c]; }
      for (int c = 0; c < K; c++) g[c] /= sum;
    }
    for (int c = 0; c < K; c++) {
      if (g[c] == 0.0) g[c] = 1e-300; else if (g[c] == 1.0) g[c] = 1.0 - 1e-16;
    }
    double mu[KT], var[KT];
    for (int c = 0; c < K; c++) { mu[c] = gen_normal(rng, d.m0, d.v0); var[c] = gen_gamma(rng, d.gshape, d.grate); }
    __syncthreads();
    for (int c = 0; c < K; c++) { P[c] = mu[c]; P[K + c] = var[c]; P[2 * K + c] = g[c]; }
    __syncthreads();
  }
};

// ===========================================================================
// C5  Beta-binomial.  P = [alpha, beta]; one factor per group (BetaBinomial.bl:17-30)
// ===========================================================================
struct BetaBinomialFam {
  struct Data {
    const int32_t* y; const int32_t* nt;   // [R][n]
    const double* cst;                     // [R][n]  lnGamma(n+1) - lnGamma(y+1) - lnGamma(n-y+1), -inf if unsupported
    int n; int chains_per_replica;
    double rate;
  };
  static constexpr int kThreads = 32;
  static constexpr bool kWarpCapable = true;
  Data d; double* P; const int32_t* y; const int32_t* nt; const double* cst; RowRange rr; int* err;
  int t0, tn;

  BPT_D static int n_reals(const Data&) { return 2; }
  BPT_D static int n_samplers(const Data&) { return 2; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; err = err_;
    const size_t rep = (size_t)(local_slot / d.chains_per_replica);
    y = d.y + rep * d.n; nt = d.nt + rep * d.n; cst = d.cst + rep * d.n;
    rr = group_rows(d.n, G, crank); t0 = threadIdx.x; tn = blockDim.x;
  }
  BPT_D void set_lanes(int lane, int width) { t0 = lane; tn = width; }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  BPT_D double fixed(int, double x, int) const { return gamma_ab(1.0, d.rate, x); }
  BPT_D void sum_groups(double a, double b, double& s, int& c) const {
    if (a <= 0.0 || b <= 0.0) {
      for (int g = rr.lo + t0; g < rr.hi; g += tn) c++;
      return;
    }
#ifndef BPT_LGAMMA
#define BPT_LGAMMA lgamma
#endif
    const double ab = BPT_LGAMMA(a + b) - BPT_LGAMMA(a) - BPT_LGAMMA(b);
    for (int g = rr.lo + t0; g < rr.hi; g += tn) {
      const double yy = (double)y[g], n = (double)nt[g];
      const double t = cst[g] + BPT_LGAMMA(yy + a) + BPT_LGAMMA(n - yy + b) + ab - BPT_LGAMMA(n + a + b);
      acc_term(t, s, c, err);
    }
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double a = k == 0 ? x : P[0], b = k == 1 ? x : P[1];
    double acc = 0.0; int cnt = 0;
    sum_groups(a, b, acc, cnt);
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    double acc = 0.0; int cnt = 0;
    sum_groups(P[0], P[1], acc, cnt);
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    logprior = gamma_ab(1.0, d.rate, P[0]) + gamma_cd(1.0, d.rate) + gamma_ab(1.0, d.rate, P[1]) + gamma_cd(1.0, d.rate);
  }
  BPT_D void forward(Rng& rng) {
    const double a = gen_exponential(rng, d.rate);
    const double b = gen_exponential(rng, d.rate);
    P[0] = a; P[1] = b;
  }
};

}  // namespace bpt
