// sllb_splitting.cu -- operator-splitting schedules (a17 / SURVEY.md section 8(f) rank 3) and the "potential
// modification" fields of the 2D2V time loop.
//
// Replaces sll_f_new_time_splitting_coeff / initialize_time_splitting_coeff
// (src/time_integration/splitting_methods/sll_m_time_splitting_coeff.F90:86-594), sll_s_compute_w_hermite
// (src/semi_lagrangian/fcisl/sll_m_fcisl.F90:413-486) and compute_jacobian
// (simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1403-1436).
//
// Every schedule of the reference is a palindrome: it is stored as its first half.  A V-stage weight may depend on
// the time step, b(dt) = b0 - 2 dt^2 c2 + 4 dt^4 c4 - 8 dt^6 c6 (the Vlasov-Poisson order-6 schemes), and the schemes
// with dim_split_V = 2 carry a second weight per V stage, dt^2 c2' (sign as tabulated), that multiplies the field of
// the modified potential.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sllb_internal.h"

using namespace sllb;

namespace {
struct StageW { double b0, c2, c4, c6; };   // weight = b0 - 2 dt^2 c2 + 4 dt^4 c4 - 8 dt^6 c6
struct Scheme {
    const char *name;      // split_case of the namelist (sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:514-554)
    int nb_split_step;     // stages per time step
    bool begin_T;
    int dim_split_V;       // 2: every V stage has a second weight, -dt^2 c2, for the field of the modified potential
    std::vector<StageW> half; // stages 1 .. (nb+1)/2
};
// The tables list the coefficient of "2 dt^2" with the sign the reference writes in front of it folded into c2, so
// that weight = b0 - 2 dt^2 c2 + ... for every scheme; the second weight of a dim_split_V = 2 stage is -dt^2 c2.
const Scheme &scheme(int k) {
    static const std::vector<Scheme> all = {
        {"SLL_STRANG_VTV", 3, false, 1, {{0.5}, {1.0}}},                                                         // :188-194
        {"SLL_STRANG_TVT", 3, true, 1, {{0.5}, {1.0}}},                                                          // :181-187
        {"SLL_LIE_TV", 2, true, 1, {{1.0}}},                                                                     // :169-174
        {"SLL_LIE_VT", 2, false, 1, {{1.0}}},                                                                    // :175-180
        {"SLL_TRIPLE_JUMP_TVT", 7, true, 1,                                                                      // :195-205
         {{0.675603595979829}, {1.351207191959658}, {-0.17560359597982855}, {-1.702414383919315}}},
        {"SLL_TRIPLE_JUMP_VTV", 7, false, 1,                                                                     // :206-216
         {{0.675603595979829}, {1.351207191959658}, {-0.17560359597982855}, {-1.702414383919315}}},
        {"SLL_ORDER6_VTV", 23, false, 1,                                                                         // :217-243
         {{0.0414649985182624}, {0.123229775946271}, {0.198128671918067}, {0.290553797799558}, {-0.0400061921041533},
          {-0.127049212625417}, {0.0752539843015807}, {-0.246331761062075}, {-0.0115113874206879}, {0.357208872795928},
          {0.23666992478693111}, {0.20477705429147008}}},
        {"SLL_ORDER6_TVT", 29, true, 1,                                                                          // :244-276
         {{0.0378593198406116}, {0.09171915262446165}, {0.102635633102435}, {0.183983170005006}, {-0.0258678882665587},
          {-0.05653436583288827}, {0.314241403071447}, {0.004914688774712854}, {-0.130144459517415}, {0.143761127168358},
          {0.106417700369543}, {0.328567693746804}, {-0.00879424312851058}, {-0.196411466486454234}, {0.20730506905689536}}},
        {"SLL_ORDER6VP_TVT", 9, true, 1,                                                                         // :278-299
         {{0.1095115577513980413559540}, {0.268722208204814693684441, 0.000805681667096178271312, 0.000017695766224036466792},
          {0.4451715080955340951457244}, {0.2312777917951853063155588, 0.003955911930042478239763, 0.000052384078562246674986},
          {-0.1093661316938642730033570}}},
        {"SLL_ORDER6VP_VTV", 9, false, 1,                                                                        // :300-323
         {{0.359950808794143627485664, -0.01359558332625151635, -8.562814848565929e-6}, {1.079852426382430882456991},
          {-0.1437147273026540434771131, -0.00385637757897273261, 0.0004883788785819335822}, {-0.579852426382430882456991},
          {0.567527837017020831982899, -0.03227361602037480885, 0.002005141087312622342}}},
        {"SLL_ORDER6VPnew_TVT", 9, true, 1,                                                                      // :324-347
         {{0.1095115577513980413559540},
          {0.268722208204814693684441, 0.000805681667096178271312, 8.643923349886021963e-6, 1.4231479258353431522e-6},
          {0.4451715080955340951457244}, {0.2312777917951853063155588, 0.003955911930042478239763, 0.000061435921436397119815},
          {-0.1093661316938642730033570}}},
        {"SLL_ORDER6VPnew1_VTV", 11, false, 1,                                                                   // :348-373
         {{0.0490864609761162454914412, 0.0000697287150553050840999}, {0.1687359505634374224481957},
          {0.2641776098889767002001462, 0.000625704827430047189169, -2.91660045768984781644e-6}, {0.377851589220928303880766},
          {0.1867359291349070543084126, 0.00221308512404532556163, 0.0000304848026170003878868, 4.98554938787506812159e-7},
          {-0.0931750795687314526579244}}},
        {"SLL_ORDER6VPnew2_VTV", 11, false, 1,                                                                   // :563-589
         {{0.083335463273305120964507, -0.00015280483587048489661, -0.0017675734111895638156, 0.00021214072262165668039},
          {0.72431592569108212422250}, {0.827694857845135145869413, -0.010726848627286273332, 0.012324362982853212700},
          {-0.4493507217041624582458844}, {-0.4110303211184402668339201, 0.014962337009932798678},
          {0.4500695920261606680467717}}},
        {"SLL_ORDER6VP2D_VTV", 11, false, 1,                                                                     // :374-396
         {{0.0490864609761162454914412, -0.00166171386175851683711044}, {0.1687359505634374224481957},
          {0.2641776098889767002001462, 0.00461492847770001641230401}, {0.377851589220928303880766},
          {0.1867359291349070543084126, -0.0000446959494108217402966857}, {-0.0931750795687314526579244}}},
        {"SLL_ORDER6VPOT_VTV", 11, false, 2,                                                                     // :397-426
         {{0.0490864609761162454914412, -0.00166171386175851683711044}, {0.1687359505634374224481957},
          {0.2641776098889767002001462, 0.00461492847770001641230401}, {0.377851589220928303880766},
          {0.1867359291349070543084126, -0.0000446959494108217402966857}, {-0.0931750795687314526579244}}},
        {"SLL_ORDER6VPOTnew1_VTV", 9, false, 2,                                                                  // :428-466
         {{0.359950808794143627485664, -0.0}, {1.079852426382430882456991}, {-0.1437147273026540434771131, -0.0139652542242388403673},
          {-0.579852426382430882456991}, {0.567527837017020831982899, -0.039247029382345626020}}},
        {"SLL_ORDER6VPOTnew2_VTV", 11, false, 2,                                                                 // :468-511
         {{0.086971698963920047813358, -1.98364114652831655458915e-6}, {0.303629319055488881944104},
          {0.560744966588102145251453, 0.00553752115152236516667268}, {0.303629319055488881944104},
          {-0.1477166655520221930648117, 0.00284218110811634663914191}, {-0.2145172762219555277764167}}},
        {"SLL_ORDER6VPOTnew3_VTV", 13, false, 2,                                                                 // :513-561
         {{0.0482332301753032567427580, -0.0002566567904012107264}, {0.2701015188126056215752542},
          {0.0482332301753032567427580, -0.0009439771580927593579}, {-0.108612186368692920020654},
          {0.2361392603742494444753990, 0.002494619878121813220}, {0.3385106675560872984454001},
          {0.3347885585502880840781703, 0.002670269183371982607658111}}},
    };
    return all[(size_t)k];
}
const int NUM_SCHEMES = 18;
} // namespace

extern "C" {

int sllb_splitting_case_from_name(const char *name, int *split_case) {
    if (!name || !split_case) return fail(SLLB_ERR_INVALID, "splitting_case_from_name: null");
    for (int k = 0; k < NUM_SCHEMES; ++k)
        if (strcmp(name, scheme(k).name) == 0) { *split_case = k; return SLLB_OK; }
    return fail(SLLB_ERR_INVALID, std::string("#split_case not defined: ") + name); // :552
}
const char *sllb_splitting_case_name(int split_case) {
    return (split_case >= 0 && split_case < NUM_SCHEMES) ? scheme(split_case).name : nullptr;
}

/* split_step(:) exactly as sll_t_splitting_coeff holds it: one entry per T stage, dim_split_V entries per V stage.
 * steps has room for SLLB_SPLIT_MAX_STEPS doubles. */
int sllb_splitting_coeff(int split_case, double dt, double *steps, int *nsteps, int *nb_split_step, int *split_begin_T,
                         int *dim_split_V) {
    if (!steps) return fail(SLLB_ERR_INVALID, "splitting_coeff: null");
    if (split_case < 0 || split_case >= NUM_SCHEMES) return fail(SLLB_ERR_INVALID, "#split_case not defined");
    const Scheme &sc = scheme(split_case);
    const int nb = sc.nb_split_step, nh = (int)sc.half.size();
    const double d2 = dt * dt, d4 = d2 * d2, d6 = d4 * d2;
    int n = 0;
    bool T = sc.begin_T;
    for (int k = 0; k < nb; ++k) {
        const StageW &w = sc.half[(size_t)(k < nh ? k : nb - 1 - k)];
        if (T) steps[n++] = w.b0;
        else {
            // the terms are added in the reference's order: b0 -/+ 2 dt^2 c2 + 4 dt^4 c4 - 8 dt^6 c6
            double v = w.b0;
            if (w.c2 != 0.0) v = v - 2.0 * d2 * w.c2;
            if (w.c4 != 0.0) v = v + 4.0 * d4 * w.c4;
            if (w.c6 != 0.0) v = v - 8.0 * d6 * w.c6;
            steps[n++] = v;
            if (sc.dim_split_V == 2) steps[n++] = -(d2 * w.c2);
        }
        T = !T;
    }
    if (nsteps) *nsteps = n;
    if (nb_split_step) *nb_split_step = nb;
    if (split_begin_T) *split_begin_T = sc.begin_T ? 1 : 0;
    if (dim_split_V) *dim_split_V = sc.dim_split_V;
    return SLLB_OK;
}

/* sll_s_compute_w_hermite: weights of the first derivative on the stencil r..s (r < 0 < s), w[k - r] */
int sllb_compute_w_hermite(int r, int s, double *w) {
    if (!w || r >= 0 || s <= 0) return fail(SLLB_ERR_INVALID, "compute_w_hermite: need r < 0 < s");
    double sum = 0.0;
    for (int i = r; i <= s; ++i) {
        if (i == 0) continue;
        double den = 1.0, num = 1.0;
        for (int j = r; j <= s; ++j) {
            if (j == i) continue;
            den *= (double)(i - j);
            if (j != 0) num *= (double)(-j);
        }
        w[i - r] = (1.0 / den) * num;
        sum += w[i - r];
    }
    w[-r] = -sum;
    return SLLB_OK;
}

} // extern "C"

namespace sllb {

// jac = (dx E1 * dy E2 - dx E2 * dy E1) * factor on the n1 x n2 periodic cells, derivatives by the stencil weights
__global__ void k_jacobian2d(const double *__restrict__ e1, const double *__restrict__ e2, int n1, int n2, int r, int s,
                             const double *__restrict__ w, double factor, double *__restrict__ jac) {
    const long long n = (long long)n1 * n2;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % n1), j = (int)(t / n1);
        double g11 = 0, g12 = 0, g21 = 0, g22 = 0;
        for (int k = r; k <= s; ++k) {
            int ii = (i + k) % n1; if (ii < 0) ii += n1;
            int jj = (j + k) % n2; if (jj < 0) jj += n2;
            const double wk = w[k - r];
            g11 += wk * e1[ii + (long long)n1 * j];
            g12 += wk * e2[ii + (long long)n1 * j];
            g21 += wk * e1[i + (long long)n1 * jj];
            g22 += wk * e2[i + (long long)n1 * jj];
        }
        jac[t] = (g11 * g22 - g12 * g21) * factor;
    }
}
// out = a * x + b * y
__global__ void k_lincomb2(const double *__restrict__ x, const double *__restrict__ y, double a, double b, long long n,
                           double *__restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        out[t] = 0.0 + x[t] * a + y[t] * b; // alpha = 0 + field(1)*step(1) + field(2)*step(2) (:1137-1141)
}
cudaError_t launch_jacobian2d(const double *e1, const double *e2, int n1, int n2, int r, int s, const double *d_w,
                              double factor, double *jac, cudaStream_t st) {
    const long long n = (long long)n1 * n2;
    k_jacobian2d<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e1, e2, n1, n2, r, s, d_w, factor, jac);
    return cudaGetLastError();
}
cudaError_t launch_lincomb2(const double *x, const double *y, double a, double b, long long n, double *out, cudaStream_t st) {
    k_lincomb2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, a, b, n, out);
    return cudaGetLastError();
}

} // namespace sllb
