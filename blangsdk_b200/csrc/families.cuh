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

enum SamplerType : int { S_REAL_SLICE = 0, S_SIMPLEX = 1, S_INT_SLICE = 3, S_INT_FIXED = 4 };   // S_INT_FIXED: CategoricalSampler / UniformSampler

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
//
// Tables for the logistic row term (tools/gen_logit_tables.py writes them, 60-digit arithmetic, rounded once).
// exp(-a) = 2^k * 2^(j/64) * 2^(f/64): m = round(-a * 64 log2 e) = 64 k + j, f = -a * 64 log2 e - m, |f| <= 1/2, so a
// degree-5 polynomial in f is exact to 4e-17 (exp_tab() of the mixture family keeps r = x - m ln2/64 by a Cody-Waite
// pair and the polynomial in r).  log(u) for u in (1, 2]: u = c_j (1 + rr) with c_j the centre of the
// j-th of 256 mantissa intervals, rr = u * fl(1/c_j) - 1, |rr| <= 1/512, log u = -log fl(1/c_j) + log1p(rr)
// (series to rr^5, truncation < 1e-17); the table holds the log of the value really multiplied by, so the
// identity is exact up to rounding.  No division, no reciprocal; 20 fp64 instructions per row (round 1: 25, with a
// Cody-Waite pair for r and the log series as rr + rr^2 q).
// {-64 log2(e), 1.5*2^52, -ln2/64 hi (21 low bits zero: m * hi exact), -ln2/64 lo}
static __constant__ double kLgtK[4] = {-92.33248261689366, 6755399441055744.0, -0.01083042469326756, -2.9815858269852933e-12};
static __constant__ double kLgtExpC[6] = {   // 1/5! .. 1/0!
    8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5, 1.0, 1.0};
static __constant__ double kLgtLogC[4] = {0.2, -0.25, 3.33333333333333333333e-01, -0.5};
// the logistic row term takes exp(-a) = 2^((m + f)/64) with f = a * (-64 log2 e) - m from ONE fma (|f| <= 1/2; the
// rounding of the constant costs a relative a * 7.7e-17 on exp(-a), < 3e-17 absolute on log1p(exp(-a))): polynomial in
// f, coefficients (ln2/64)^i / i!, i = 5 .. 0 (tools/gen_logit_tables.py)
static __constant__ double kLgtExp2C[6] = {1.2417843701716925e-12, 5.732851688640402e-10, 2.1173137155464776e-07,
                                           5.86490495505617e-05, 0.010830424696249145, 1.0};
static __constant__ double kLgtExpTabC[64] = {
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
static __constant__ double2 kLgtLogTabC[256] = {
    {0.9980506822612085, 0.0019512201312618048}, {0.9941747572815534, 0.0058422756242284025},
    {0.9903288201160542, 0.00971824946892134}, {0.9865125240847784, 0.013579258126380866},
    {0.982725527831094, 0.01742541671385917}, {0.9789674952198852, 0.021256839025415152},
    {0.9752380952380952, 0.025073637552115963}, {0.9715370018975332, 0.02887592350185452},
    {0.9678638941398866, 0.03266380681879157}, {0.9642184557438794, 0.03643739620243109},
    {0.9606003752345216, 0.04019679912633676}, {0.9570093457943926, 0.04394212185649872},
    {0.9534450651769087, 0.04767346946935691}, {0.9499072356215214, 0.051390945869489356},
    {0.9463955637707948, 0.05509465380697375}, {0.9429097605893186, 0.058784694894427614},
    {0.9394495412844037, 0.06246116962373623}, {0.9360146252285192, 0.06612417738247342},
    {0.9326047358834244, 0.06977381647002284}, {0.9292196007259528, 0.07341018411340675},
    {0.9258589511754068, 0.07703337648282704}, {0.9225225225225225, 0.08064348870692671},
    {0.9192100538599641, 0.0842406148877762}, {0.9159212880143113, 0.08782484811559137},
    {0.9126559714795008, 0.09139628048318858}, {0.9094138543516874, 0.09495500310018257},
    {0.9061946902654867, 0.09850110610693318}, {0.9029982363315696, 0.10203467868824434},
    {0.8998242530755711, 0.1055558090868232}, {0.8966725043782837, 0.10906458461650242},
    {0.893542757417103, 0.11256109167523176}, {0.8904347826086957, 0.11604541575784262},
    {0.8873483535528596, 0.11951764146859183}, {0.8842832469775475, 0.1229778525334875},
    {0.8812392426850258, 0.12642613181240342}, {0.8782161234991424, 0.12986256131098461},
    {0.8752136752136752, 0.1332872221923487}, {0.8722316865417377, 0.13670019478858866},
    {0.8692699490662139, 0.14010155861207893}, {0.8663282571912013, 0.14349139236659045},
    {0.863406408094435, 0.1468697739582177}, {0.8605042016806723, 0.1502367805061219},
    {0.8576214405360134, 0.15359248835309428}, {0.8547579298831386, 0.15693697307594154},
    {0.8519134775374376, 0.16027030949569976}, {0.8490878938640133, 0.16359257168767763},
    {0.8462809917355372, 0.1669038329913337}, {0.8434925864909391, 0.17020416601999047},
    {0.8407224958949097, 0.17349364267038928}, {0.8379705400981997, 0.17677233413208748},
    {0.835236541598695, 0.18004031089670358}, {0.832520325203252, 0.18329764276701008},
    {0.8298217179902755, 0.1865443988658801}, {0.827140549273021, 0.18978064764508826},
    {0.8244766505636071, 0.19300645689397097}, {0.8218298555377207, 0.19622189374794538},
    {0.8192, 0.19942702469689366}, {0.8165869218500797, 0.20262191559341292},
    {0.8139904610492846, 0.2058066316609327}, {0.8114104595879557, 0.20898123750170533},
    {0.8088467614533965, 0.2121457971046684}, {0.8062992125984252, 0.21530037385318387},
    {0.8037676609105181, 0.21844503053265552}, {0.8012519561815337, 0.221579829338027},
    {0.7987519500780031, 0.22470483188116222}, {0.7962674961119751, 0.2278200991981116},
    {0.7937984496124031, 0.2309256917562647}, {0.7913446676970634, 0.2340216694613928},
    {0.7889060092449923, 0.2371080916645822}, {0.7864823348694316, 0.24018501716906146},
    {0.7840735068912711, 0.24325250423692324}, {0.7816793893129771, 0.24631061059574413},
    {0.7792998477929984, 0.24935939344510277}, {0.7769347496206374, 0.2523989094629994},
    {0.7745839636913767, 0.25542921481217845}, {0.7722473604826546, 0.25845036514635467},
    {0.7699248120300752, 0.2614624156163463}, {0.767616191904048, 0.26446542087611596},
    {0.7653213751868461, 0.2674594350887206}, {0.7630402384500745, 0.270444511932174},
    {0.7607726597325408, 0.27342070460522006}, {0.7585185185185185, 0.2763880658330221},
    {0.7562776957163959, 0.27934664787276714}, {0.7540500736377025, 0.28229650251918836},
    {0.7518355359765051, 0.28523768111000464}, {0.7496339677891655, 0.28817023453128227},
    {0.7474452554744525, 0.29109421322271756}, {0.745269286754003, 0.2940096671828415},
    {0.7431059506531205, 0.2969166459741508}, {0.7409551374819102, 0.2998151987281622},
    {0.7388167388167388, 0.30270537415039545}, {0.7366906474820144, 0.3055872205252843},
    {0.7345767575322812, 0.30846078572101604}, {0.7324749642346209, 0.3113261171943025},
    {0.7303851640513552, 0.3141832619950823}, {0.7283072546230441, 0.3170322667711571},
    {0.7262411347517731, 0.3198731777727608}, {0.7241867043847242, 0.32270604085706495},
    {0.7221438645980254, 0.3255309014926197}, {0.720112517580872, 0.3283478047637331},
    {0.7180925666199158, 0.3311567953747882}, {0.7160839160839161, 0.3339579176544999},
    {0.7140864714086471, 0.3367512155601126}, {0.7121001390820584, 0.33953673268153894},
    {0.710124826629681, 0.3423145122454413}, {0.7081604426002767, 0.3450845971192569},
    {0.7062068965517241, 0.3478470298151671}, {0.7042640990371389, 0.35060185249401155},
    {0.7023319615912208, 0.35334910696915045}, {0.7004103967168263, 0.35608883471027075},
    {0.6984993178717599, 0.3588210768471437}, {0.6965986394557823, 0.36154587417332895},
    {0.694708276797829, 0.3642632671498289}, {0.6928281461434371, 0.36697329590869393},
    {0.6909581646423751, 0.36967600025657915}, {0.6890982503364738, 0.3723714196782514},
    {0.687248322147651, 0.3750595933400518}, {0.6854082998661312, 0.3777405600933096},
    {0.6835781041388518, 0.3804143584777117}, {0.681757656458056, 0.3830810267246269},
    {0.6799468791500664, 0.38574060276038585}, {0.6781456953642384, 0.38839312420951694},
    {0.6763540290620872, 0.39103862839794096}, {0.6745718050065876, 0.3936771523561221},
    {0.6727989487516426, 0.39630873282217804}, {0.6710353866317169, 0.3989334062449492},
    {0.6692810457516339, 0.4015512087870281}, {0.6675358539765319, 0.4041621763277484},
    {0.6657997399219766, 0.4067663444661362}, {0.6640726329442282, 0.40936374852382174},
    {0.6623544631306598, 0.4119544235479141}, {0.6606451612903226, 0.4145384043138392},
    {0.6589446589446589, 0.4171157253281397}, {0.6572528883183568, 0.4196864208312405},
    {0.6555697823303457, 0.4222505248001782}, {0.6538952745849298, 0.42480807095129525},
    {0.6522292993630573, 0.4273590927429007}, {0.650571791613723, 0.4299036233778954},
    {0.6489226869455006, 0.43244169580636654}, {0.6472819216182049, 0.43497334272814603},
    {0.6456494325346784, 0.4374985965953402}, {0.6440251572327044, 0.4400174896148242},
    {0.6424090338770388, 0.4425300537507073}, {0.6408010012515645, 0.44503632072676685},
    {0.6392009987515606, 0.4475363220288514}, {0.6376089663760897, 0.4500300889072539},
    {0.6360248447204969, 0.45251765237905556}, {0.6344485749690211, 0.454999043230441},
    {0.6328800988875154, 0.457474292018984}, {0.6313193588162762, 0.45994342907590513},
    {0.6297662976629766, 0.4624064845083028}, {0.6282208588957056, 0.46486348820135487},
    {0.6266829865361077, 0.4673144698204952}, {0.6251526251526252, 0.4697594588135616},
    {0.6236297198538368, 0.4721984844129205}, {0.6221142162818954, 0.47463157563756214},
    {0.6206060606060606, 0.4770587612951732}, {0.619105199516324, 0.4794800699841836},
    {0.617611580217129, 0.4818955300957873}, {0.6161251504211793, 0.48430516981594035},
    {0.6146458583433373, 0.486709017127335}, {0.6131736526946108, 0.4891070998113477},
    {0.6117084826762246, 0.4914994454499676}, {0.6102502979737783, 0.49388608142769846},
    {0.6087990487514863, 0.4962670349334404}, {0.6073546856465006, 0.4986423329623476},
    {0.6059171597633136, 0.5010120023176661}, {0.6044864226682408, 0.5033760696125467},
    {0.6030624263839811, 0.5057345612718396}, {0.6016451233842538, 0.5080875035338663},
    {0.6002344665885111, 0.5104349224521714}, {0.5988304093567252, 0.5127768438972524},
    {0.5974329054842473, 0.5151132935582721}, {0.5960419091967404, 0.5174442969447476},
    {0.59465737514518, 0.519769879388223}, {0.593279258400927, 0.5220900660439202},
    {0.5919075144508671, 0.5244048818923714}, {0.5905420991926182, 0.526714351741034},
    {0.5891829689298044, 0.5290185002258843}, {0.5878300803673938, 0.531317351812995},
    {0.586483390607102, 0.5336109308000944}, {0.5851428571428572, 0.5358992613181066},
    {0.5838084378563284, 0.5381823673326752}, {0.5824800910125142, 0.5404602726456693},
    {0.5811577752553916, 0.5427330008966718}, {0.579841449603624, 0.5450005755644521},
    {0.5785310734463277, 0.5472630199684217}, {0.5772266065388951, 0.5495203572700718},
    {0.5759280089988752, 0.5517726104743967}, {0.574635241301908, 0.5540198024313014},
    {0.5733482642777156, 0.5562619558369912}, {0.5720670391061452, 0.5584990932353476},
    {0.5707915273132664, 0.5607312370192884}, {0.5695216907675195, 0.5629584094321125},
    {0.5682574916759157, 0.5651806325688301}, {0.5669988925802879, 0.5673979283774777},
    {0.5657458563535912, 0.5696103186604182}, {0.5644983461962514, 0.5718178250756288},
    {0.5632563256325632, 0.5740204691379711}, {0.562019758507135, 0.5762182722204505},
    {0.56078860898138, 0.5784112555554607}, {0.5595628415300546, 0.5805994402360135},
    {0.5583424209378408, 0.5827828472169571}, {0.5571273122959739, 0.5849614973161792},
    {0.5559174809989142, 0.587135411215799}, {0.5547128927410617, 0.5893046094633444},
    {0.5535135135135135, 0.5914691124729174}, {0.552319309600863, 0.5936289405263474},
    {0.5511302475780409, 0.5957841137743308}, {0.5499462943071965, 0.5979346522375593},
    {0.5487674169346195, 0.600080575807836}, {0.5475935828877005, 0.6022219042491792},
    {0.5464247598719317, 0.6043586571989144}, {0.5452609158679447, 0.606490854168755},
    {0.5441020191285866, 0.6086185145458719}, {0.542948038176034, 0.6107416575939496},
    {0.5417989417989418, 0.612860302454235}, {0.5406546990496304, 0.6149744681465705},
    {0.5395152792413066, 0.6170841735704201}, {0.5383806519453207, 0.6191894375058825},
    {0.5372507869884575, 0.6212902786146943}, {0.5361256544502618, 0.6233867154412224},
    {0.5350052246603971, 0.6254787664134464}, {0.5338894681960376, 0.6275664498439304},
    {0.5327783558792925, 0.6296497839307846}, {0.5316718587746625, 0.6317287867586178},
    {0.5305699481865285, 0.6338034762994782}, {0.5294725956566702, 0.6358738704137865},
    {0.5283797729618163, 0.6379399868512585}, {0.5272914521112255, 0.6400018432518172},
    {0.526207605344296, 0.6420594571464973}, {0.5251282051282051, 0.6441128459583394},
    {0.5240532241555783, 0.646162027003275}, {0.5229826353421859, 0.6482070174910025},
    {0.5219164118246687, 0.6502478345258552}, {0.5208545269582909, 0.6522844951076588},
    {0.5197969543147208, 0.6543170161325811}, {0.5187436676798379, 0.6563454143939738},
    {0.5176946410515673, 0.6583697065832043}, {0.5166498486377397, 0.6603899092904802},
    {0.5156092648539778, 0.6624060390056649}, {0.5145728643216081, 0.6644181121190849},
    {0.5135406218655968, 0.6664261449223305}, {0.5125125125125125, 0.6684301536090458},
    {0.5114885114885115, 0.6704301542757128}, {0.5104685942173479, 0.6724261629224279},
    {0.5094527363184079, 0.6744181954536684}, {0.5084409136047666, 0.6764062676790545},
    {0.5074331020812686, 0.6783903953141012}, {0.506429277942631, 0.6803705939809637},
    {0.5054294175715696, 0.6823468792091757}, {0.5044334975369458, 0.68431926643638},
    {0.5034414945919371, 0.6862877710090521}, {0.5024533856722276, 0.6882524081832171},
    {0.5014691478942214, 0.6902131931251577}, {0.5004887585532747, 0.6921701409121187}};

// the tables live in shared memory (lanes index them with different j); one copy per CTA, filled once
struct LogitTables { double2 lg[256]; double ex[64]; };
__shared__ LogitTables s_logit_tab;
BPT_D void logit_table_init() {   // call before the first __syncthreads() of the kernel
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_logit_tab.lg[i] = kLgtLogTabC[i];
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
    asm("ld.shared.f64 %0, [%1];" : "=d"(te[w]) : "r"(tab + 4096u + ((uint32_t)(m[w] & 63) << 3)));
    r[w] = fma(a, kLgtK[0], -kd[w]);        // f = -a * 64 log2 e - m, one rounding
    p[w] = kLgtExp2C[0];
  }
#pragma unroll
  for (int i = 1; i < 6; i++)
#pragma unroll
    for (int w = 0; w < W; w++) p[w] = fma(p[w], r[w], kLgtExp2C[i]);
  double q[W], rr[W], lc[W];
#pragma unroll
  for (int w = 0; w < W; w++) {
    // t = exp(-|z|) = 2^k * (2^(j/64) * p) in (0, 1];  u = 1 + t in (1, 2]
    const double tp = te[w] * p[w];
    const double t = __hiloint2double(__double2hiint(tp) + ((m[w] >> 6) << 20), __double2loint(tp));
    const double u = 1.0 + t;
    const unsigned j = min(255u, (unsigned)(__double2hiint(u) - 0x3ff00000) >> 12);
    double2 cj;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cj.x), "=d"(cj.y) : "r"(tab + (j << 4)));
    rr[w] = fma(u, cj.x, -1.0);
    lc[w] = cj.y;
    q[w] = kLgtLogC[0];
  }
#pragma unroll
  for (int i = 1; i < 4; i++)
#pragma unroll
    for (int w = 0; w < W; w++) q[w] = fma(q[w], rr[w], kLgtLogC[i]);
#pragma unroll
  for (int w = 0; w < W; w++) {
    const double l = fma(rr[w], fma(rr[w], q[w], 1.0), lc[w]);    // log1p(exp(-|z|)) = log c_j + rr (1 + rr q)
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
// table-driven fp64 exp / log for the mixture kernel (operand ranges known, see MixtureFam::obs_term)
// ---------------------------------------------------------------------------
// The logistic kernel's tables (logit_table_init()): exp in 10 and log in 10 fp64 instructions (a degree-11
// polynomial exp and a reciprocal + atanh-series log took 17 and 25).  `tab` = logit_table_address().
BPT_D double exp_tab(double x, uint32_t tab) {   // exp(x) for -700 <= x <= 700 (caller clamps)
  double kd = fma(x, -kLgtK[0], kLgtK[1]);       // low word = m = round(x * 64 log2 e) = 64 k + j
  const int m = __double2loint(kd);
  kd -= kLgtK[1];
  double te;
  asm("ld.shared.f64 %0, [%1];" : "=d"(te) : "r"(tab + 4096u + ((uint32_t)(m & 63) << 3)));
  double r = fma(kd, kLgtK[2], x);
  r = fma(kd, kLgtK[3], r);
  double p = kLgtExpC[0];
#pragma unroll
  for (int i = 1; i < 6; i++) p = fma(p, r, kLgtExpC[i]);
  const double tp = te * p;
  return __hiloint2double(__double2hiint(tp) + ((m >> 6) << 20), __double2loint(tp));
}
BPT_D double log_tab(double a, uint32_t tab) {   // log(a) for normal positive a
  const int hi = __double2hiint(a);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));   // [1, 2)
  double2 cj;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cj.x), "=d"(cj.y) : "r"(tab + (((uint32_t)hi & 0x000ff000u) >> 8)));
  const double rr = fma(m, cj.x, -1.0);
  double q = kLgtLogC[0];
#pragma unroll
  for (int i = 1; i < 4; i++) q = fma(q, rr, kLgtLogC[i]);
  const double lm = fma(rr * rr, q, rr) + cj.y;
  return fma((double)e, 6.93147180559945309417e-01, lm);
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
// Family 6: i.i.d. observations from a law with 1-3 latent parameters, each with its own prior (SURVEY 8(f)-4:
// more of R/distributions/*.bl).  Every law of the distribution is one factor PER OBSERVATION, also the laws that
// read only parameters (Poisson's -mean, Student's -log sigma ...): n of them, and n out-of-support counts.
// ===========================================================================
enum {
  IID_POISSON = 1, IID_BINOMIAL, IID_NEGBINOMIAL, IID_STUDENTT, IID_GEOMETRIC, IID_LAPLACE, IID_EXPONENTIAL, IID_GAMMA,
  IID_WEIBULL, IID_GUMBEL, IID_LOGISTICDIST, IID_HALFSTUDENTT, IID_LOGLOGISTIC, IID_GOMPERTZ, IID_YULESIMON,
  IID_UNIFORM_MAX,
  IID_N_DISTS
};
enum { PRIOR_NORMAL = 0, PRIOR_GAMMA = 1, PRIOR_EXPONENTIAL = 2, PRIOR_BETA = 3, PRIOR_UNIFORM = 4 };

// Per distribution: number of parameters, number of laws (`logf` blocks, file order), and per law which
// parameters its signature names (bit j = parameter j) and whether it names the realization.
struct IidShape { int n_params, n_laws; unsigned char mask[4], reads_y[4]; };
BPT_HD IidShape iid_shape(int dist) {
  switch (dist) {
    case IID_POISSON:      return {1, 3, {1, 1, 0, 0}, {1, 0, 1, 0}};   // mean
    case IID_BINOMIAL:     return {1, 2, {1, 0, 0, 0}, {1, 1, 0, 0}};   // p (trials are data)
    case IID_NEGBINOMIAL:  return {2, 2, {1, 3, 0, 0}, {1, 1, 0, 0}};   // r, p
    case IID_STUDENTT:     return {3, 4, {0, 4, 1, 7}, {0, 0, 0, 1}};   // nu, mu, sigma
    case IID_GEOMETRIC:    return {1, 1, {1, 0, 0, 0}, {1, 0, 0, 0}};   // p
    case IID_LAPLACE:      return {2, 1, {3, 0, 0, 0}, {1, 0, 0, 0}};   // location, scale
    case IID_EXPONENTIAL:  return {1, 4, {1, 1, 0, 1}, {1, 1, 0, 0}};   // rate (-> Gamma(1.0, rate))
    case IID_GAMMA:        return {2, 4, {3, 2, 1, 2}, {1, 1, 0, 0}};   // shape, rate
    case IID_WEIBULL:      return {2, 3, {3, 3, 3, 0}, {0, 1, 1, 0}};   // scale, shape
    case IID_GUMBEL:       return {2, 2, {3, 3, 0, 0}, {0, 1, 0, 0}};   // location, scale
    case IID_LOGISTICDIST: return {2, 3, {2, 3, 3, 0}, {0, 1, 1, 0}};   // location, scale
    case IID_HALFSTUDENTT: return {2, 3, {1, 2, 3, 0}, {0, 0, 1, 0}};   // nu, sigma
    case IID_LOGLOGISTIC:  return {2, 3, {3, 3, 3, 0}, {0, 1, 1, 0}};   // scale, shape
    case IID_GOMPERTZ:     return {2, 2, {3, 3, 0, 0}, {0, 1, 0, 0}};   // shape, scale
    case IID_YULESIMON:    return {1, 1, {1, 0, 0, 0}, {1, 0, 0, 0}};   // rho
    case IID_UNIFORM_MAX:  return {1, 2, {1, 1, 0, 0}, {0, 1, 0, 0}};   // max (min = 0.0: F/Doomsday.bl:9)
  }
  return {0, 0, {0, 0, 0, 0}, {0, 0, 0, 0}};
}
BPT_D double iid_gamma_law(int law, double shape, double rate, double y) {   // Gamma.bl:14-31
  switch (law) {
    case 0: return (shape <= 0.0 || rate <= 0.0 || y <= 0.0) ? BPT_NEG_INF : (shape - 1.0) * log(y * rate);
    case 1: return (rate <= 0.0 || y <= 0.0) ? BPT_NEG_INF : -y * rate;
    case 2: return shape <= 0.0 ? BPT_NEG_INF : -lgamma(shape);
    default: return rate <= 0.0 ? BPT_NEG_INF : log(rate);
  }
}
// one `logf` block of R/distributions/<Name>.bl; th = parameters, y = realization, x = trials (Binomial)
BPT_D double iid_law(int dist, int law, const double* th, double y, double x) {
  const double ninf = BPT_NEG_INF;
  switch (dist) {
    case IID_POISSON: {                                   // Poisson.bl:11-26
      const double mean = th[0];
      if (law == 0) return (mean <= 0.0 || y < 0.0) ? ninf : y * log(mean);
      if (law == 1) return mean <= 0.0 ? ninf : -mean;
      return y < 0.0 ? ninf : -lgamma(y + 1.0);
    }
    case IID_BINOMIAL: {                                  // Binomial.bl:14-27
      const double p = th[0];
      if (y < 0.0 || x <= 0.0 || y > x) return ninf;
      if (law == 0) return (p < 0.0 || p > 1.0) ? ninf : y * log(p) + (x - y) * log(1.0 - p);
      return lgamma(x + 1.0) - lgamma(y + 1.0) - lgamma(x - y + 1.0);
    }
    case IID_NEGBINOMIAL: {                               // NegativeBinomial.bl:14-28
      const double r = th[0], p = th[1];
      if (r <= 0.0 || y < 0.0) return ninf;
      if (law == 0) {
        const double nn = y + r - 1.0;
        const double t = lgamma(nn + 1.0) - lgamma(y + 1.0) - lgamma(nn - y + 1.0);
        return t != t ? ninf : t;
      }
      return (p <= 0.0 || p >= 1.0) ? ninf : y * log(p) + r * log(1.0 - p);
    }
    case IID_STUDENTT: {                                  // StudentT.bl:13-33
      const double nu = th[0], mu = th[1], sigma = th[2];
      if (law == 0) return -0.5 * log(3.14159265358979323846);
      if (law == 1) return sigma <= 0.0 ? ninf : -log(sigma);
      if (law == 2) return nu <= 0.0 ? ninf : lgamma((nu + 1.0) / 2.0) - 0.5 * log(nu) - lgamma(nu / 2.0);
      if (nu <= 0.0 || sigma <= 0.0) return ninf;
      const double dy = y - mu;
      return -((nu + 1.0) / 2.0) * log(1.0 + (1.0 / nu) * (1.0 / (sigma * sigma)) * (dy * dy));
    }
    case IID_GEOMETRIC: {                                 // Geometric.bl:10-16
      const double p = th[0];
      return (p <= 0.0 || p >= 1.0 || y < 0.0) ? ninf : y * log(1.0 - p) + log(p);
    }
    case IID_LAPLACE: {                                   // Laplace.bl:12-17
      const double loc = th[0], scale = th[1];
      return scale <= 0.0 ? ninf : -log(2.0 * scale) - fabs(y - loc) / scale;
    }
    case IID_EXPONENTIAL: return iid_gamma_law(law, 1.0, th[0], y);   // Exponential.bl:11
    case IID_GAMMA: return iid_gamma_law(law, th[0], th[1], y);
    case IID_WEIBULL: {                                   // Weibull.bl:14-33
      const double scale = th[0], shape = th[1];
      if (scale <= 0.0 || shape <= 0.0) return ninf;
      if (law == 0) return log(shape) - (shape * log(scale));
      if (y <= 0.0) return ninf;
      return law == 1 ? (shape - 1.0) * log(y) : -pow(y / scale, shape);
    }
    case IID_GUMBEL: {                                    // Gumbel.bl:13-22
      const double loc = th[0], scale = th[1];
      if (scale <= 0.0) return ninf;
      if (law == 0) return -log(scale);
      const double z = (loc - y) / scale;
      return -exp(z) + z;
    }
    case IID_LOGISTICDIST: {                              // Logistic.bl:13-26
      const double loc = th[0], scale = th[1];
      if (scale <= 0.0) return ninf;
      if (law == 0) return -log(scale);
      const double z = (loc - y) / scale;
      return law == 1 ? z : -2.0 * log(1.0 + exp(z));
    }
    case IID_HALFSTUDENTT: {                              // HalfStudentT.bl:12-28
      const double nu = th[0], sigma = th[1];
      if (law == 0) return nu <= 0.0 ? ninf : log(2.0) + lgamma((nu + 1.0) / 2.0) - lgamma(nu / 2.0) - 0.5 * log(nu * 3.14159265358979323846);
      if (law == 1) return sigma <= 0.0 ? ninf : -log(sigma);
      if (y < 0.0 || sigma <= 0.0 || nu <= 0.0) return ninf;
      const double q = y / sigma;
      return -((nu + 1.0) / 2.0) * log(1.0 + 1.0 / nu * (q * q));
    }
    case IID_LOGLOGISTIC: {                               // LogLogistic.bl:14-33
      const double scale = th[0], shape = th[1];
      if (law == 0) return (scale <= 0.0 || shape <= 0.0) ? ninf : log(shape) - (shape * log(scale));
      if (y < 0.0 || scale <= 0.0 || shape <= 0.0) return ninf;
      return law == 1 ? shape * log(y) - log(y) : -2.0 * log(1.0 + pow(y / scale, shape));
    }
    case IID_GOMPERTZ: {                                  // Gompertz.bl:13-25
      const double shape = th[0], scale = th[1];
      if (law == 0) return (shape <= 0.0 || scale <= 0.0) ? ninf : log(shape / scale);
      if (y < 0.0 || scale <= 0.0 || shape <= 0.0) return ninf;
      return (y / scale) - (shape * (exp(y / scale) - 1.0));
    }
    case IID_UNIFORM_MAX: {                               // ContinuousUniform.bl:14-23 with min = 0.0
      const double mn = 0.0, mx = th[0];
      if (law == 0) return (mx - mn <= 0.0) ? ninf : -log(mx - mn);
      return (mn <= y && y <= mx) ? 0.0 : ninf;
    }
    default: {                                            // YuleSimon.bl:10-16
      const double rho = th[0];
      if (rho <= 0.0 || y < 0.0) return ninf;
      return log(rho) + (lgamma(1.0 + y) + lgamma(rho + 1.0) - lgamma(1.0 + y + (rho + 1.0)));
    }
  }
}

struct IidFam {
  struct Data {
    const double* y; const double* trials; int n, dist, n_params;
    int prior_kind[3]; double prior_a[3], prior_b[3];
  };
  static constexpr int kThreads = 512;          // launch bound; small data sets run 128 threads per CTA
  static int launch_threads(const Data& d) { return d.n >= 32768 ? 512 : 128; }
  static constexpr bool kWarpCapable = true;
  Data d; double* P; RowRange rr; int* err;
  int t0, tn;   // this thread's index in the chain group and the group size

  BPT_D static int n_reals(const Data& d) { return d.n_params; }
  BPT_D static int n_samplers(const Data& d) { return d.n_params; }
  BPT_D void init(const Data& d_, int, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; rr = group_rows(d.n, G, crank); err = err_; t0 = threadIdx.x; tn = blockDim.x;
  }
  BPT_D void set_lanes(int lane, int width) { t0 = lane; tn = width; }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}

  // the prior's factors that read the realization
  BPT_D double fixed(int k, double x, int) const {
    const double a = d.prior_a[k], b = d.prior_b[k];
    switch (d.prior_kind[k]) {
      case PRIOR_NORMAL: return normal_f2(a, b, x);
      case PRIOR_GAMMA: return gamma_ab(a, b, x);
      case PRIOR_EXPONENTIAL: return gamma_ab(1.0, a, x);          // Exponential.bl:11 -> Gamma(1.0, rate)
      case PRIOR_BETA:                                                // Beta.bl:14-27
        if (x <= 0.0 || x >= 1.0 || a <= 0.0 || b <= 0.0) return BPT_NEG_INF;
        return (a - 1.0) * log(x) + (b - 1.0) * log1p(-x);
      default: return (a <= x && x <= b) ? 0.0 : BPT_NEG_INF;       // ContinuousUniform.bl:16-19
    }
  }
  // the prior's constant factors
  BPT_D double fixed_constants(int k) const {
    const double a = d.prior_a[k], b = d.prior_b[k];
    switch (d.prior_kind[k]) {
      case PRIOR_NORMAL: return -0.5 * BPT_LOG_2PI + (b <= 0.0 ? BPT_NEG_INF : -0.5 * log(b));
      case PRIOR_GAMMA: return gamma_cd(a, b);
      case PRIOR_EXPONENTIAL: return gamma_cd(1.0, a);
      case PRIOR_BETA: return (a <= 0.0 || b <= 0.0) ? BPT_NEG_INF : lgamma(a + b) - lgamma(a) - lgamma(b);
      default: return (b - a <= 0.0) ? BPT_NEG_INF : -log(b - a);
    }
  }
  // Sum and out-of-support count of the annealed factors connected to parameter k (k < 0: of all of them),
  // at parameters th.  Laws that do not read the realization are evaluated once and counted n times.
  BPT_D void factors(int k, const double* th, GroupReducer& red, double& s, int& c) const {
    const int n = d.n;
    const IidShape sh = iid_shape(d.dist);
    double acc = 0.0, par = 0.0; int cnt = 0, pc = 0;
    for (int l = 0; l < sh.n_laws; l++) {
      if (k >= 0 && !((sh.mask[l] >> k) & 1)) continue;
      if (!sh.reads_y[l]) {
        const double v = iid_law(d.dist, l, th, 0.0, 0.0);
        if (v > BPT_NEG_INF) par += n * v;
        else if (v != v) atomicOr(err, ERR_NAN);
        else pc += n;
      } else {
        int i = rr.lo + t0;
        for (; i + 3 * tn < rr.hi; i += 4 * tn) {      // four rows' loads in flight, same summation order
          double yv[4], xv[4];
#pragma unroll
          for (int u = 0; u < 4; u++) { yv[u] = d.y[i + u * tn]; xv[u] = d.trials ? d.trials[i + u * tn] : 0.0; }
#pragma unroll
          for (int u = 0; u < 4; u++) acc_term(iid_law(d.dist, l, th, yv[u], xv[u]), acc, cnt, err);
        }
        for (; i < rr.hi; i += tn)
          acc_term(iid_law(d.dist, l, th, d.y[i], d.trials ? d.trials[i] : 0.0), acc, cnt, err);
      }
    }
    red.reduce(acc, cnt);
    s = acc + par; c = cnt + pc;
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    double th[3] = {P[0], d.n_params > 1 ? P[1] : 0.0, d.n_params > 2 ? P[2] : 0.0};
    th[k] = x;
    factors(k, th, red, s, c);
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    const double th[3] = {P[0], d.n_params > 1 ? P[1] : 0.0, d.n_params > 2 ? P[2] : 0.0};
    factors(-1, th, red, loglik, noos);
    double lp = 0.0;
    for (int j = 0; j < d.n_params; j++) lp += fixed_constants(j) + fixed(j, P[j], 0);
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {   // declaration order
    for (int j = 0; j < d.n_params; j++) {
      const double a = d.prior_a[j], b = d.prior_b[j];
      double v;
      switch (d.prior_kind[j]) {
        case PRIOR_NORMAL: v = gen_normal(rng, a, b); break;
        case PRIOR_GAMMA: v = gen_gamma(rng, a, b); break;
        case PRIOR_EXPONENTIAL: v = gen_exponential(rng, a); break;
        case PRIOR_BETA: v = gen_beta(rng, a, b); break;
        default: v = gen_uniform(rng, a, b); break;
      }
      P[j] = v;
    }
  }
};

// ===========================================================================
// Family 8: F/LinRegression.bl generalised to p covariates.  P = [w_0 .. w_{p-1}, sigma]
//   w_j ~ Normal(0, v0), sigma ~ ContinuousUniform(0, smax), y_i ~ Normal(sum_j w_j x_ij, sigma * sigma)
// Same layout as the logistic family: column-major X and a per-chain cache in HBM, here of the residuals
// res_i = mean_i - y_i ("mean - x" of Normal.bl:21-24), so a move of w_k streams one column and the cache.
// ===========================================================================
struct LinRegFam {
  struct Data {
    const double* xt;      // [p][npad] column-major design matrix
    const double* y;       // [npad]
    double* res;           // [localSlots][npad]
    int n, npad, p;
    double v0, smax;
  };
  static constexpr int kThreads = 512;
  static constexpr bool kWarpCapable = false;
  Data d; double* P; double* res; RowRange rr; int* err;

  BPT_D static int n_reals(const Data& d) { return d.p + 1; }
  BPT_D static int n_samplers(const Data& d) { return d.p + 1; }
  BPT_D void init(const Data& d_, int local_slot, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; res = d.res + (size_t)local_slot * d.npad; rr = group_rows(d.n, G, crank); err = err_;
  }
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  BPT_D double fixed(int k, double x, int) const {
    if (k < d.p) return normal_f2(0.0, d.v0, x);
    return (0.0 <= x && x <= d.smax) ? 0.0 : BPT_NEG_INF;         // ContinuousUniform.bl:16-19
  }
  // this CTA's share of sum_i (res_i + delta * X[i][k])^2; eight steps' loads are issued before their FMAs (the
  // pass is two loads and four FMAs per pair: latency-, not bandwidth-bound)
  BPT_D double sum_sq(int k, double delta) const {
    const double* col = d.xt + (size_t)k * d.npad;
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    const double2* pr = reinterpret_cast<const double2*>(res + rr.lo);
    const double2* px = reinterpret_cast<const double2*>(col + rr.lo);
    const int T = blockDim.x;
    double s0 = 0.0, s1 = 0.0;
    int q = threadIdx.x;
    for (; q + 7 * T < n_pairs; q += 8 * T) {
      double2 r[8], xv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { r[u] = pr[q + u * T]; xv[u] = px[q + u * T]; }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double a = fma(delta, xv[u].x, r[u].x), b = fma(delta, xv[u].y, r[u].y);
        s0 = fma(a, a, s0); s1 = fma(b, b, s1);
      }
    }
    for (; q < n_pairs; q += T) {
      const double2 r = pr[q], xv = px[q];
      const double a = fma(delta, xv.x, r.x), b = fma(delta, xv.y, r.y);
      s0 = fma(a, a, s0); s1 = fma(b, b, s1);
    }
    double s = s0 + s1;
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) { const int i = rr.hi - 1; const double a = fma(delta, col[i], res[i]); s = fma(a, a, s); }
    return s;
  }
  // w_k: the n third blocks of Normal.bl; sigma: the n second and the n third blocks (Normal.bl:17-24)
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const double sigma = k == d.p ? x : P[d.p], v = sigma * sigma;
    if (v <= 0.0) { s = 0.0; c = (k == d.p ? 2 : 1) * d.n; return; }
    double q = k < d.p ? sum_sq(k, x - P[k]) : sum_sq(0, 0.0); int z = 0;
    red.reduce(q, z);
    s = -0.5 * q / v;
    if (k == d.p) s += d.n * (-0.5 * log(v));
    c = 0;
  }
  BPT_D void commit(int k, double x, int) {
    if (k < d.p) {
      const double delta = x - P[k];
      const double* col = d.xt + (size_t)k * d.npad;
      if (delta != 0.0)
        for (int i = rr.lo + 2 * threadIdx.x; i < rr.hi; i += 2 * blockDim.x) {
          if (i + 1 < rr.hi) {
            double2 r = *reinterpret_cast<const double2*>(res + i);
            const double2 xv = *reinterpret_cast<const double2*>(col + i);
            r.x = fma(delta, xv.x, r.x); r.y = fma(delta, xv.y, r.y);
            *reinterpret_cast<double2*>(res + i) = r;
          } else res[i] = fma(delta, col[i], res[i]);
        }
    }
    __syncthreads();   // every thread has read P[k] for its delta before anyone overwrites it
    P[k] = x;
  }
  // rebuild the residuals from scratch: sum_j w_j * X[i][j], j ascending from 0.0, minus y_i
  BPT_D void refresh() {
    for (int i = rr.lo + threadIdx.x; i < rr.hi; i += blockDim.x) {
      double m = 0.0;
      for (int j = 0; j < d.p; j++) m = fma(P[j], d.xt[(size_t)j * d.npad + i], m);
      res[i] = m - d.y[i];
    }
    __syncthreads();   // sum_sq reads pairs written by other threads
  }
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) {
    refresh();
    const double sigma = P[d.p], v = sigma * sigma;
    loglik = d.n * (-0.5 * BPT_LOG_2PI);
    if (v <= 0.0) noos = 2 * d.n;
    else {
      double q = sum_sq(0, 0.0); int z = 0;
      red.reduce(q, z);
      loglik += d.n * (-0.5 * log(v)) + (-0.5 * q / v);
      noos = 0;
    }
    double lp = 0.0;
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int j = 0; j < d.p; j++) lp += f01 + normal_f2(0.0, d.v0, P[j]);
    lp += (d.smax <= 0.0 ? BPT_NEG_INF : -log(d.smax)) + fixed(d.p, sigma, 0);
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {
    for (int j = 0; j < d.p; j++) { const double w = gen_normal(rng, 0.0, d.v0); P[j] = w; }
    const double s = gen_uniform(rng, 0.0, d.smax); P[d.p] = s;
  }
};

// ===========================================================================
// Family 7: F/HierarchicalModel.bl, not marginalised (the rocket-launch example of the Blang documentation)
//   a ~ Exponential(rate_a), b ~ Exponential(rate_b), p_g | a, b ~ Beta(a, b), y_g | p_g ~ Binomial(n_g, p_g)
// P = [a, b, p_0 .. p_{G-1}], one real slice sampler each.  The Beta blocks are fixed factors (p_g is latent); only
// the Binomial blocks are annealed, and a move of p_g touches one of them: no reduction, every thread of the chain
// group computes the same scalars.  Small by nature (G <= 62): one warp per chain, many replicas per launch.
// ===========================================================================
struct HierFam {
  struct Data { const int* y; const int* n; int G; double rate_a, rate_b; };
  static constexpr int kThreads = 32;
  static constexpr bool kWarpCapable = true;
  static constexpr int kFusedParams = kMaxParams;   // a, b and up to 62 group probabilities per chain
  Data d; double* P; int* err;
  double sum_log_p, sum_log1m_p; bool p_inside;   // over the groups, refreshed when a sampler execution begins

  BPT_D static int n_reals(const Data& d) { return d.G + 2; }
  BPT_D static int n_samplers(const Data& d) { return d.G + 2; }
  BPT_D void init(const Data& d_, int, double* P_, int, int, int* err_) {
    d = d_; P = P_; err = err_; sum_log_p = 0.0; sum_log1m_p = 0.0; p_inside = false;   // set by begin_execution()
  }
  BPT_D void set_lanes(int, int) {}
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D static void tables_init() {}
  BPT_D void sums() {
    double s1 = 0.0, s2 = 0.0; bool ok = true;
    for (int g = 0; g < d.G; g++) {
      const double p = P[2 + g];
      if (p <= 0.0 || p >= 1.0) ok = false; else { s1 += log(p); s2 += log1p(-p); }
    }
    sum_log_p = s1; sum_log1m_p = s2; p_inside = ok;
  }
  BPT_D void begin_execution() { sums(); }
  // Beta.bl:14-41 with latent parameters, the blocks that name variable k
  BPT_D double fixed(int k, double x, int) const {
    const double a = k == 0 ? x : P[0], b = k == 1 ? x : P[1];
    if (k == 0) {        // Exponential prior + G x { (a-1) log p_g, lnGamma(a+b), -lnGamma(a) }
      if (a <= 0.0 || !p_inside || b <= 0.0) return BPT_NEG_INF;
      return gamma_ab(1.0, d.rate_a, a) + (a - 1.0) * sum_log_p + d.G * (lgamma(a + b) - lgamma(a));
    }
    if (k == 1) {        // Exponential prior + G x { (b-1) log1p(-p_g), lnGamma(a+b), -lnGamma(b) }
      if (b <= 0.0 || !p_inside || a <= 0.0) return BPT_NEG_INF;
      return gamma_ab(1.0, d.rate_b, b) + (b - 1.0) * sum_log1m_p + d.G * (lgamma(a + b) - lgamma(b));
    }
    if (x <= 0.0 || x >= 1.0 || a <= 0.0 || b <= 0.0) return BPT_NEG_INF;
    return (a - 1.0) * log(x) + (b - 1.0) * log1p(-x);
  }
  BPT_D static double binomial_block(int law, double y, double n, double p) {
    const double th[1] = {p};
    return iid_law(IID_BINOMIAL, law, th, y, n);
  }
  // annealed factors connected to variable k: none for a and b, the group's first Binomial block for p_g
  BPT_D void annealed(int k, double x, int, GroupReducer&, double& s, int& c) const {
    s = 0.0; c = 0;
    if (k < 2) return;
    const int g = k - 2;
    acc_term(binomial_block(0, (double)d.y[g], (double)d.n[g], x), s, c, err);
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer&, double& loglik, int& noos, double& logprior) {
    double s = 0.0; int c = 0;
    for (int g = 0; g < d.G; g++) {
      acc_term(binomial_block(0, (double)d.y[g], (double)d.n[g], P[2 + g]), s, c, err);
      acc_term(binomial_block(1, (double)d.y[g], (double)d.n[g], P[2 + g]), s, c, err);
    }
    loglik = s; noos = c;
    const double a = P[0], b = P[1];
    double lp = gamma_ab(1.0, d.rate_a, a) + gamma_cd(1.0, d.rate_a) + gamma_ab(1.0, d.rate_b, b) + gamma_cd(1.0, d.rate_b);
    for (int g = 0; g < d.G; g++) {
      const double p = P[2 + g];
      const bool in = p > 0.0 && p < 1.0;
      lp += (in && a > 0.0) ? (a - 1.0) * log(p) : BPT_NEG_INF;
      lp += (in && b > 0.0) ? (b - 1.0) * log1p(-p) : BPT_NEG_INF;
      lp += (a > 0.0 && b > 0.0) ? lgamma(a + b) : BPT_NEG_INF;
      lp += a > 0.0 ? -lgamma(a) : BPT_NEG_INF;
      lp += b > 0.0 ? -lgamma(b) : BPT_NEG_INF;
    }
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {   // topological order: a, b, then every p_g
    const double a = gen_exponential(rng, d.rate_a);
    const double b = gen_exponential(rng, d.rate_b);
    P[0] = a; P[1] = b;
    for (int g = 0; g < d.G; g++) { const double p = gen_beta(rng, a, b); P[2 + g] = p; }
  }
};

// ===========================================================================
// Family 9: F/SimpleHierarchicalModel.bl -- family 7 with LATENT launch counts, updated by the integer slice
// sampler with stepping out (IntSliceSampler.java:53-132):
//   a, b ~ Exponential; p_g | a, b ~ Beta(a + 0.5, b + 0.5); L_g ~ Poisson(mean); failures_g ~ Binomial(1 + L_g, p_g)
// P = [a, b, p_0 .. p_{G-1}, L_0 .. L_{G-1}]; the integers are held as exact doubles (G <= 31).
// ===========================================================================
struct RocketsLatentFam {
  struct Data { const int* y; int G; double rate_a, rate_b, pois_mean; };
  static constexpr int kThreads = 32;
  static constexpr bool kWarpCapable = true;
  static constexpr bool kHasIntSlice = true;
  static constexpr int kFusedParams = kMaxParams;
  Data d; double* P; int* err;
  double sum_log_p, sum_log1m_p; bool p_inside;   // over the groups, refreshed when a sampler execution begins

  BPT_D static int n_reals(const Data& d) { return 2 * d.G + 2; }
  BPT_D static int n_samplers(const Data& d) { return 2 * d.G + 2; }
  BPT_D void init(const Data& d_, int, double* P_, int, int, int* err_) {
    d = d_; P = P_; err = err_; sum_log_p = 0.0; sum_log1m_p = 0.0; p_inside = false;   // set by begin_execution()
  }
  BPT_D void set_lanes(int, int) {}
  BPT_D int sampler_type(int k) const { return k < d.G + 2 ? S_REAL_SLICE : S_INT_SLICE; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D static void tables_init() {}
  BPT_D void begin_execution() {
    double s1 = 0.0, s2 = 0.0; bool ok = true;
    for (int g = 0; g < d.G; g++) {
      const double p = P[2 + g];
      if (p <= 0.0 || p >= 1.0) ok = false; else { s1 += log(p); s2 += log1p(-p); }
    }
    sum_log_p = s1; sum_log1m_p = s2; p_inside = ok;
  }
  BPT_D double poisson_block(int law, double L) const { const double th[1] = {d.pois_mean}; return iid_law(IID_POISSON, law, th, L, 0.0); }
  BPT_D static double binomial_block(int law, double y, double n, double p) { const double th[1] = {p}; return iid_law(IID_BINOMIAL, law, th, y, n); }
  BPT_D double fixed(int k, double x, int) const {
    const int G = d.G;
    if (k >= G + 2) return poisson_block(0, x) + poisson_block(2, x);        // Poisson.bl: the blocks that name the realization
    const double a = (k == 0 ? x : P[0]), b = (k == 1 ? x : P[1]);
    const double al = a + 0.5, be = b + 0.5;                                   // Beta(a + 0.5, b + 0.5)
    if (k == 0) {
      if (al <= 0.0 || be <= 0.0 || !p_inside) return BPT_NEG_INF;
      return gamma_ab(1.0, d.rate_a, a) + (al - 1.0) * sum_log_p + G * (lgamma(al + be) - lgamma(al));
    }
    if (k == 1) {
      if (al <= 0.0 || be <= 0.0 || !p_inside) return BPT_NEG_INF;
      return gamma_ab(1.0, d.rate_b, b) + (be - 1.0) * sum_log1m_p + G * (lgamma(al + be) - lgamma(be));
    }
    if (x <= 0.0 || x >= 1.0 || al <= 0.0 || be <= 0.0) return BPT_NEG_INF;
    return (al - 1.0) * log(x) + (be - 1.0) * log1p(-x);
  }
  // annealed: p_g sees the group's first Binomial block; L_g sees both (the number of trials is 1 + L_g)
  BPT_D void annealed(int k, double x, int, GroupReducer&, double& s, int& c) const {
    s = 0.0; c = 0;
    const int G = d.G;
    if (k < 2) return;
    if (k < G + 2) {
      const int g = k - 2;
      acc_term(binomial_block(0, (double)d.y[g], 1.0 + P[G + 2 + g], x), s, c, err);
    } else {
      const int g = k - (G + 2);
      acc_term(binomial_block(0, (double)d.y[g], 1.0 + x, P[2 + g]), s, c, err);
      acc_term(binomial_block(1, (double)d.y[g], 1.0 + x, P[2 + g]), s, c, err);
    }
  }
  BPT_D void commit(int k, double x, int) { P[k] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer&, double& loglik, int& noos, double& logprior) {
    const int G = d.G;
    double s = 0.0; int c = 0;
    for (int g = 0; g < G; g++) {
      acc_term(binomial_block(0, (double)d.y[g], 1.0 + P[G + 2 + g], P[2 + g]), s, c, err);
      acc_term(binomial_block(1, (double)d.y[g], 1.0 + P[G + 2 + g], P[2 + g]), s, c, err);
    }
    loglik = s; noos = c;
    const double a = P[0], b = P[1], al = a + 0.5, be = b + 0.5;
    double lp = gamma_ab(1.0, d.rate_a, a) + gamma_cd(1.0, d.rate_a) + gamma_ab(1.0, d.rate_b, b) + gamma_cd(1.0, d.rate_b);
    for (int g = 0; g < G; g++) {
      const double p = P[2 + g];
      const bool in = p > 0.0 && p < 1.0;
      lp += (in && al > 0.0) ? (al - 1.0) * log(p) : BPT_NEG_INF;
      lp += (in && be > 0.0) ? (be - 1.0) * log1p(-p) : BPT_NEG_INF;
      lp += (al > 0.0 && be > 0.0) ? lgamma(al + be) : BPT_NEG_INF;
      lp += al > 0.0 ? -lgamma(al) : BPT_NEG_INF;
      lp += be > 0.0 ? -lgamma(be) : BPT_NEG_INF;
      lp += poisson_block(0, P[G + 2 + g]) + poisson_block(1, 0.0) + poisson_block(2, P[G + 2 + g]);
    }
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {   // topological order: a, b, then per group p_g and L_g
    const double a = gen_exponential(rng, d.rate_a);
    const double b = gen_exponential(rng, d.rate_b);
    P[0] = a; P[1] = b;
    for (int g = 0; g < d.G; g++) {
      const double p = gen_beta(rng, a + 0.5, b + 0.5);
      const int L = gen_poisson_small(rng, d.pois_mean);
      P[2 + g] = p; P[d.G + 2 + g] = (double)L;
    }
  }
};

// ===========================================================================
// C2  Bayesian logistic regression.  P = w[0..p)
//     per-chain cache eta_i = s_i * (x_i . w) in HBM, s_i = +1 if y_i = 1 else -1 (the rows of
//     the device copy of X are sign-folded once at set_model, which is exact); a move of w_k
//     needs only column k of X:  eta_i(x) = eta_i + (x - w_k) * s_i X[i][k]
// ===========================================================================
template <int THREADS, int MINB = 1, int L2AHEAD = 0>
struct LogisticFamT {
  static constexpr int kMinBlocks = MINB;   // CTAs per SM the register budget must allow
  // L2AHEAD > 0: the per-chain cache does not fit the L2 (many chains per GPU) and streams from HBM; each step
  // also asks the L2 for the cache lines L2AHEAD steps ahead, so the register prefetch (one step) only ever
  // waits for an L2 hit
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
      if (L2AHEAD > 0 && (threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(pe + (1 + L2AHEAD) * THREADS));
      logit_pair<true>(fma(delta, xa.x, ea.x), fma(delta, xa.y, ea.y), py, acc, cnt, err, tab);
      pe += THREADS; px += THREADS; py += 2 * THREADS;
      if (whole < 32) break;
      whole -= THREADS;
      if (whole >= 32) { ea = pe[THREADS]; xa = px[THREADS]; }
      if (L2AHEAD > 0 && (threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(pe + (1 + L2AHEAD) * THREADS));
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
  BPT_D static void tables_init() { logit_table_init(); }
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
                      int& c, uint32_t tab) const {
    double acc = 0.0;
#pragma unroll
    for (int q0 = 0; q0 < KT; q0 += 4) {     // four interleaved exp chains at a time
      double arg[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const double dd = mu[q0 + q] - yv; arg[q] = fmax(fma(b[q0 + q], dd * dd, a[q0 + q]), -700.0); }
#pragma unroll
      for (int q = 0; q < 4; q++) acc = fma(pi[q0 + q], exp_tab(arg[q], tab), acc);
    }
    if (acc > 1e-250 && acc < 1e250) s += log_tab(acc, tab);
    else obs_slow(yv, mu, a, b, pi, s, c);
  }
  // not inlined: the loop gets its own register allocation instead of competing with the sampler state
  __device__ __noinline__ void sum_obs(double& s, int& c) const {
    const uint32_t tab = logit_table_address();
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
      obs_term(v.x, mu, a, b, pi, s, c, tab);
      obs_term(v.y, mu, a, b, pi, s, c, tab);
    }
    if (((rr.hi - rr.lo) & 1) && threadIdx.x == 0) obs_term(d.y[rr.hi - 1], mu, a, b, pi, s, c, tab);
  }
  // First evaluation of a sampler execution (always at the current state): every component is evaluated, and
  // per observation the sum R of the components the sampler does NOT move is stored (plus, for the simplex
  // sampler, the two moved components' unweighted densities).  j0 / j1 = moved components (j1 = -1 if one).
  __device__ __noinline__ void pass_first(int j0, int j1, double& s, int& c) const {
    const uint32_t tab = logit_table_address();
    double mu[KT], a[KT], b[KT], pi[KT];
#pragma unroll
    for (int q = 0; q < KT; q++) { mu[q] = Cst[q]; a[q] = Cst[KT + q]; b[q] = Cst[2 * KT + q]; pi[q] = Cst[3 * KT + q]; }
    // lean evaluation of one row, no side effects: R (unmoved components), S (all), the moved components' densities
    auto lean = [&](double yv, double& R, double& S, double& E1, double& E2) {
      double M = 0.0; R = 0.0; E1 = 0.0; E2 = 0.0;
#pragma unroll
      for (int q0 = 0; q0 < KT; q0 += 4) {
        double arg[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const double dd = mu[q0 + q] - yv; arg[q] = fmax(fma(b[q0 + q], dd * dd, a[q0 + q]), -700.0); }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double e = exp_tab(arg[q], tab);
          const bool moved = (q0 + q == j0) || (q0 + q == j1);
          if (q0 + q == j0) E1 = e;
          if (q0 + q == j1) E2 = e;
          if (moved) M = fma(pi[q0 + q], e, M); else R = fma(pi[q0 + q], e, R);
        }
      }
      S = R + M;
    };
    auto one = [&](double yv, double& r_out, double& e1_out, double& e2_out) {
      double R, S;
      lean(yv, R, S, e1_out, e2_out);
      if (S > 1e-250 && S < 1e250) { s += log_tab(S, tab); r_out = R; }
      else { obs_slow(yv, mu, a, b, pi, s, c); r_out = CUDART_NAN; }   // NaN marks a row the cached passes redo literally
    };
    const int n_pairs = (rr.hi - rr.lo) >> 1;
    for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
      const int i = rr.lo + 2 * p;
      const double2 v = *reinterpret_cast<const double2*>(d.y + i);
      double2 r, x1, x2;
      // both rows as straight-line code; one warp vote decides whether all lanes keep the lean logs
      double S0, S1;
      lean(v.x, r.x, S0, x1.x, x2.x); lean(v.y, r.y, S1, x1.y, x2.y);
      const double l0 = log_tab(S0, tab), l1 = log_tab(S1, tab);
      const bool ok = S0 > 1e-250 && S0 < 1e250 && S1 > 1e-250 && S1 < 1e250;
      if (__all_sync(__activemask(), ok)) { s += l0; s += l1; }
      else { one(v.x, r.x, x1.x, x2.x); one(v.y, r.y, x1.y, x2.y); }
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
    const uint32_t tab = logit_table_address();
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
      if (!SIMPLEX) { const double dd = mu0 - yv; S = fma(pi0, exp_tab(fmax(fma(b0, dd * dd, a0), -700.0), tab), R); }
      else S = fma(pi0, E1, fma(pi1, E2, R));
      if (S > 1e-250 && S < 1e250) s += log_tab(S, tab);
      else slow(yv);
    };
    // 2*kB rows as straight-line code (their exp/log chains interleave); one warp vote decides whether every
    // lane may keep the lean results, otherwise the rows are redone one by one (same values, same order)
    auto group = [&](const double2* v, const double2* r, const double2* x1, const double2* x2, int n_valid) {
      double S[2 * kB], l[2 * kB];
      bool ok = true;
#pragma unroll
      for (int u = 0; u < kB; u++) {
        const int w = SIMPLEX ? u : 0;
        const double yv[2] = {v[u].x, v[u].y}, R[2] = {r[u].x, r[u].y};
        const double E1[2] = {x1[w].x, x1[w].y}, E2[2] = {x2[w].x, x2[w].y};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (!SIMPLEX) { const double dd = mu0 - yv[h]; S[2 * u + h] = fma(pi0, exp_tab(fmax(fma(b0, dd * dd, a0), -700.0), tab), R[h]); }
          else S[2 * u + h] = fma(pi0, E1[h], fma(pi1, E2[h], R[h]));
        }
      }
#pragma unroll
      for (int i = 0; i < 2 * kB; i++) {
        l[i] = log_tab(S[i], tab);
        ok = ok && ((S[i] > 1e-250 && S[i] < 1e250) || (i >> 1) >= n_valid);
      }
      if (__all_sync(__activemask(), ok)) {
#pragma unroll
        for (int i = 0; i < 2 * kB; i++) if ((i >> 1) < n_valid) s += l[i];
        return;
      }
#pragma unroll
      for (int u = 0; u < kB; u++)
        if (u < n_valid) {
          const int w = SIMPLEX ? u : 0;
          one(v[u].x, r[u].x, x1[w].x, x2[w].x); one(v[u].y, r[u].y, x1[w].y, x2[w].y);
        }
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
      const int n_valid = min(kB, (n_pairs - cur + stride - 1) / stride);   // pairs of this batch that exist
      group(v, r, x1, x2, n_valid);
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
    else if (j1 < 0) pass_cached<false, 2>(j0, j1, acc, cnt);
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
// Family 10: blang.examples.MixtureModel (F/MixtureModel.bl) with K components and the cluster indicators NOT
// marginalised.  P = [mu[K], var[K], pi[K], z[n] as exact doubles]
//   pi ~ Dirichlet(conc); mu_k ~ Normal(m0, v0); var_k ~ Gamma(shape, rate); z_i | pi ~ Categorical(pi) (latent);
//   y_i | z_i ~ Normal(means.get(z_i), variances.get(z_i)) (observed: three annealed factors each).
// Samplers: k < K mu_k, K <= k < 2K var_k, 2K the simplex, 2K + 1 + i the indicator z_i, moved by CategoricalSampler =
// IntSliceSampler with the fixed window [0, K) (R/mcmc/CategoricalSampler.xtend:24-28); with K > 10 the window test
// IntSliceSampler.accept (:115-132) runs.  Connections are the DSL's static ones: the Normal's parameter closures read
// the whole lists `means` / `variances`, so every mean's sampler sees every f2, every variance's sampler every f1 and
// f2; the simplex sampler sees the Dirichlet factor and the n Categorical factors log pi[z_i].
// A small-model family (3K + n <= 96 values, 2K + 1 + n <= 65 samplers): one CTA per chain.
// ===========================================================================
struct MixIndFam {
  struct Data { const double* y; int n, K; double m0, v0, gshape, grate, conc; };
  static constexpr int kThreads = 128;
  static constexpr bool kWarpCapable = false;
  static constexpr bool kHasIntFixed = true;
  Data d; double* P; int* err; int t0, tn;

  BPT_D static int n_reals(const Data& d) { return 3 * d.K + d.n; }
  BPT_D static int n_samplers(const Data& d) { return 2 * d.K + 1 + d.n; }
  BPT_D void init(const Data& d_, int, double* P_, int, int, int* err_) { d = d_; P = P_; err = err_; t0 = threadIdx.x; tn = blockDim.x; }
  BPT_D int sampler_type(int k) const { return k < 2 * d.K ? S_REAL_SLICE : (k == 2 * d.K ? S_SIMPLEX : S_INT_FIXED); }
  BPT_D int sampler_var(int k) const { return k <= 2 * d.K ? k : 3 * d.K + (k - 2 * d.K - 1); }
  BPT_D int simplex_dim() const { return d.K; }
  BPT_D int int_lo(int) const { return 0; }
  BPT_D int int_hi(int) const { return d.K; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  // candidate-substituted weight of component c for the simplex sampler (SimplexWritableVariable.set :26-30)
  BPT_D double pi_at(int c, double x, int aux) const {
    const int K = d.K, nx = aux == K - 1 ? 0 : aux + 1;
    if (c == aux) return x;
    if (c == nx) return fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x);
    return P[2 * K + c];
  }
  BPT_D double fixed(int k, double x, int aux) const {
    const int K = d.K;
    if (k < K) return normal_f2(d.m0, d.v0, x);
    if (k < 2 * K) return gamma_ab(d.gshape, d.grate, x);
    if (k == 2 * K) {     // Dirichlet.bl:13-22, then the n Categorical factors (Categorical.bl:12-15), connection order
      double sum = 0.0;
      for (int c = 0; c < K; c++) { if (d.conc < 0.0) return BPT_NEG_INF; sum += (d.conc - 1.0) * log(pi_at(c, x, aux)); }
      for (int i = 0; i < d.n; i++) sum += log(pi_at((int)P[3 * K + i], x, aux));
      return sum;
    }
    const int z = (int)x;   // indicator i: its own Categorical factor
    return (z < 0 || z >= K) ? BPT_NEG_INF : log(P[2 * K + z]);
  }
  // Normal.bl:17-24 of observation i with component z
  BPT_D void obs_f1(double var, double& s, int& c) const { acc_term(var <= 0.0 ? BPT_NEG_INF : -0.5 * log(var), s, c, err); }
  BPT_D void obs_f2(double mu, double var, double yv, double& s, int& c) const {
    const double dd = mu - yv;
    acc_term(var <= 0.0 ? BPT_NEG_INF : -0.5 * (dd * dd) / var, s, c, err);
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    const int K = d.K;
    double acc = 0.0; int cnt = 0;
    if (k < 2 * K) {
      for (int i = t0; i < d.n; i += tn) {
        const int z = (int)P[3 * K + i];
        const double mu = (k < K && z == k) ? x : P[z], var = (k >= K && z == k - K) ? x : P[K + z];
        if (k >= K) obs_f1(var, acc, cnt);
        obs_f2(mu, var, d.y[i], acc, cnt);
      }
    } else if (k > 2 * K) {
      const int i = k - 2 * K - 1, z = (int)x;
      if (t0 == 0) {
        if (z < 0 || z >= K) cnt += 2;
        else { obs_f1(P[K + z], acc, cnt); obs_f2(P[z], P[K + z], d.y[i], acc, cnt); }
      }
    }
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int aux) {
    const int K = d.K;
    const int nx = aux == K - 1 ? 0 : aux + 1;
    const double comp = k == 2 * K ? fmax(0.0, (P[2 * K + aux] + P[2 * K + nx]) - x) : 0.0;
    __syncthreads();
    if (k == 2 * K) { P[2 * K + aux] = x; P[2 * K + nx] = comp; }
    else P[sampler_var(k)] = x;
    __syncthreads();
  }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    const int K = d.K;
    double acc = 0.0; int cnt = 0;
    for (int i = t0; i < d.n; i += tn) {
      const int z = (int)P[3 * K + i];
      acc += -0.5 * BPT_LOG_2PI;
      obs_f1(P[K + z], acc, cnt); obs_f2(P[z], P[K + z], d.y[i], acc, cnt);
    }
    red.reduce(acc, cnt);
    loglik = acc; noos = cnt;
    double lp = 0.0;
    for (int c = 0; c < K; c++) lp += (d.conc - 1.0) * log(P[2 * K + c]);
    lp += d.conc < 0.0 ? BPT_NEG_INF : (-K * lgamma(d.conc) + lgamma(K * d.conc));
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    for (int c = 0; c < K; c++) {
      lp += f01 + normal_f2(d.m0, d.v0, P[c]);
      lp += gamma_ab(d.gshape, d.grate, P[K + c]) + gamma_cd(d.gshape, d.grate);
    }
    for (int i = 0; i < d.n; i++) lp += log(P[2 * K + (int)P[3 * K + i]]);
    logprior = lp;
  }
  BPT_D void forward(Rng& rng) {   // pi, then (mu_k, var_k) per component, then the indicators (GraphAnalysis.createForwardSimulator order)
    const int K = d.K;
    double g[16]; double sum = 0.0;
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
    for (int c = 0; c < K; c++) { if (g[c] == 0.0) g[c] = 1e-300; else if (g[c] == 1.0) g[c] = 1.0 - 1e-16; }
    double mu[16], var[16];
    for (int c = 0; c < K; c++) { mu[c] = gen_normal(rng, d.m0, d.v0); var[c] = gen_gamma(rng, d.gshape, d.grate); }
    __syncthreads();
    for (int c = 0; c < K; c++) { P[c] = mu[c]; P[K + c] = var[c]; P[2 * K + c] = g[c]; }
    for (int i = 0; i < d.n; i++) {   // Generators.categorical :163 -> first index whose cumulative sum exceeds U
      const double u = rng.next_double();
      double cum = 0.0; int z = K - 1;
      for (int c = 0; c < K; c++) { cum += g[c]; if (u < cum) { z = c; break; } }
      P[3 * K + i] = (double)z;
    }
    __syncthreads();
  }
};

// ===========================================================================
// Family 12: F/CustomAnnealRef.bl / F/CustomAnnealTest.bl, the pair of T/TestEndToEnd.xtend:169-186:
//   mu ~ Normal(m0, v0); x (observed) | mu ~ Normal(mu, s2)
// annealed automatically (three exponentiated factors) or "manually" (one LogPotential whose body multiplies
// Normal::distribution(mu, s2).logDensity(x) by an AnnealingParameter: SampledModel.sumOtherAnnealed :219-225).
// A custom-annealed factor that is LINEAR in the annealing parameter is carried as pre-annealing log-likelihood: the
// annealed density fixed + beta * L is the same number at every beta either way (logDensity :161-176 adds the same
// terms in another order), so samplers, swaps and SCM weights take identical decisions.  What differs in the
// reference is only what it REPORTS for such a factor: energy 0 and no out-of-support softening (exponent == null);
// the engine reports -L.  P = [mu].
// ===========================================================================
struct AnnealPairFam {
  struct Data { double x, m0, v0, s2; int manual; };
  static constexpr int kThreads = 32;
  static constexpr bool kWarpCapable = true;
  Data d; double* P; int* err;
  BPT_D static int n_reals(const Data&) { return 1; }
  BPT_D static int n_samplers(const Data&) { return 1; }
  BPT_D void init(const Data& d_, int, double* P_, int, int, int* err_) { d = d_; P = P_; err = err_; }
  BPT_D void set_lanes(int, int) {}
  BPT_D int sampler_type(int) const { return S_REAL_SLICE; }
  BPT_D int sampler_var(int) const { return 0; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  BPT_D double fixed(int, double x, int) const { return normal_f2(d.m0, d.v0, x); }
  // Normal.bl:14-24, the three blocks in file order (what Normal::distribution(mu, s2).logDensity(x) adds up)
  BPT_D void blocks(double mu, double& s, int& c) const {
    s = 0.0; c = 0;
    acc_term(-0.5 * BPT_LOG_2PI, s, c, err);
    acc_term(d.s2 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.s2), s, c, err);
    const double dd = mu - d.x;
    acc_term(d.s2 <= 0.0 ? BPT_NEG_INF : -0.5 * (dd * dd) / d.s2, s, c, err);
  }
  BPT_D void annealed(int, double x, int, GroupReducer&, double& s, int& c) const {
    // the sampler of mu sees the block that names the mean (automatic) or the whole LogPotential (manual)
    if (d.manual) { blocks(x, s, c); return; }
    const double dd = x - d.x;      // Normal.bl:21-24 with mean = mu
    s = 0.0; c = 0;
    acc_term(d.s2 <= 0.0 ? BPT_NEG_INF : -0.5 * (dd * dd) / d.s2, s, c, err);
  }
  BPT_D void commit(int, double x, int) { P[0] = x; }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer&, double& loglik, int& noos, double& logprior) const {
    blocks(P[0], loglik, noos);
    logprior = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0)) + normal_f2(d.m0, d.v0, P[0]);
  }
  BPT_D void forward(Rng& rng) { const double mu = gen_normal(rng, d.m0, d.v0); P[0] = mu; }
};

// ===========================================================================
// Family 11: change point, the smallest model with a DiscreteUniform latent (UniformSampler =
// IntSliceSampler with the fixed window [minInclusive, maxExclusive), R/mcmc/UniformSampler.xtend:21-27).
//   mu_0, mu_1 ~ Normal(m0, v0); c ~ DiscreteUniform(0, n + 1); y_i | mu, c ~ Normal(i < c ? mu_0 : mu_1, s2), s2 constant.
// P = [mu_0, mu_1, c as an exact double]; all three samplers see every observation's f2 (the mean closure reads all three).
// ===========================================================================
struct ChangePointFam {
  struct Data { const double* y; int n; double m0, v0, s2; };
  static constexpr int kThreads = 512;
  static int launch_threads(const Data& d) { return d.n >= 32768 ? 512 : 128; }
  static constexpr bool kWarpCapable = false;
  static constexpr bool kHasIntFixed = true;
  Data d; double* P; RowRange rr; int* err; int t0, tn;

  BPT_D static int n_reals(const Data&) { return 3; }
  BPT_D static int n_samplers(const Data&) { return 3; }
  BPT_D void init(const Data& d_, int, double* P_, int G, int crank, int* err_) {
    d = d_; P = P_; rr = group_rows(d.n, G, crank); err = err_; t0 = threadIdx.x; tn = blockDim.x;
  }
  BPT_D int sampler_type(int k) const { return k < 2 ? S_REAL_SLICE : S_INT_FIXED; }
  BPT_D int sampler_var(int k) const { return k; }
  BPT_D int simplex_dim() const { return 0; }
  BPT_D int int_lo(int) const { return 0; }
  BPT_D int int_hi(int) const { return d.n + 1; }
  BPT_D void begin_execution() {}
  BPT_D static void tables_init() {}
  BPT_D double fixed(int k, double x, int) const {
    if (k < 2) return normal_f2(d.m0, d.v0, x);
    return (0.0 <= x && x < (double)(d.n + 1)) ? 0.0 : BPT_NEG_INF;      // DiscreteUniform.bl:16-20
  }
  BPT_D void sum_f2(double mu0, double mu1, double cp, double& s, int& c) const {
    if (d.s2 <= 0.0) { c += rr.hi - rr.lo > 0 && t0 == 0 ? rr.hi - rr.lo : 0; return; }
    for (int i = rr.lo + t0; i < rr.hi; i += tn) {
      const double dd = ((double)i < cp ? mu0 : mu1) - d.y[i];
      s += -0.5 * (dd * dd) / d.s2;
    }
  }
  BPT_D void annealed(int k, double x, int, GroupReducer& red, double& s, int& c) const {
    double acc = 0.0; int cnt = 0;
    sum_f2(k == 0 ? x : P[0], k == 1 ? x : P[1], k == 2 ? x : P[2], acc, cnt);
    red.reduce(acc, cnt);
    s = acc; c = cnt;
  }
  BPT_D void commit(int k, double x, int) { __syncthreads(); P[k] = x; __syncthreads(); }
  BPT_D void refresh() {}
  BPT_D void full(GroupReducer& red, double& loglik, int& noos, double& logprior) const {
    double acc = 0.0; int cnt = 0;
    sum_f2(P[0], P[1], P[2], acc, cnt);
    red.reduce(acc, cnt);
    loglik = acc + d.n * (-0.5 * BPT_LOG_2PI);
    if (d.s2 <= 0.0) cnt += d.n; else loglik += d.n * (-0.5 * log(d.s2));   // the n f1 factors
    noos = cnt;
    const double f01 = -0.5 * BPT_LOG_2PI + (d.v0 <= 0.0 ? BPT_NEG_INF : -0.5 * log(d.v0));
    logprior = 2.0 * f01 + normal_f2(d.m0, d.v0, P[0]) + normal_f2(d.m0, d.v0, P[1]) - log((double)(d.n + 1)) + fixed(2, P[2], 0);
  }
  BPT_D void forward(Rng& rng) {
    const double a = gen_normal(rng, d.m0, d.v0), b = gen_normal(rng, d.m0, d.v0);
    const int cp = rng.next_int(d.n + 1);          // Generators.discreteUniform :205-222
    __syncthreads();
    P[0] = a; P[1] = b; P[2] = (double)cp;
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
    const double ab = lgamma(a + b) - lgamma(a) - lgamma(b);
    for (int g = rr.lo + t0; g < rr.hi; g += tn) {
      const double yy = (double)y[g], n = (double)nt[g];
      const double t = cst[g] + lgamma(yy + a) + lgamma(n - yy + b) + ab - lgamma(n + a + b);
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
