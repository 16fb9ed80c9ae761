/*
 * sll_oracle_split.c -- CPU restatement of the operator-splitting tables and of the "potential modification"
 * of the 2D2V time loop.  TEST INFRASTRUCTURE ONLY, like sll_oracle.c.
 *
 * Restates src/time_integration/splitting_methods/sll_m_time_splitting_coeff.F90:157-594 (every split_case the
 * 2D2V simulation's namelist accepts, simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/
 * sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:514-554), sll_s_compute_w_hermite
 * (src/semi_lagrangian/fcisl/sll_m_fcisl.F90:413-486) and compute_jacobian (...poisson_serial.F90:1403-1436).
 * The coefficients are data of the published schemes (Strang; Yoshida triple jump; Blanes-Moan O6-11/O6-14;
 * Casas-Crouseilles-Faou-Mehrenberger Vlasov-Poisson order-6 schemes with dt-dependent weights).
 *
 * Case numbering used by the oracle and by include/sll_b200.h (SLLB_SPLIT_*): 0 STRANG_VTV, 1 STRANG_TVT, 2 LIE_TV,
 * 3 LIE_VT, 4 TRIPLE_JUMP_TVT, 5 TRIPLE_JUMP_VTV, 6 ORDER6_VTV, 7 ORDER6_TVT, 8 ORDER6VP_TVT, 9 ORDER6VP_VTV,
 * 10 ORDER6VPnew_TVT, 11 ORDER6VPnew1_VTV, 12 ORDER6VPnew2_VTV, 13 ORDER6VP2D_VTV, 14 ORDER6VPOT_VTV,
 * 15 ORDER6VPOTnew1_VTV, 16 ORDER6VPOTnew2_VTV, 17 ORDER6VPOTnew3_VTV.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void mirror(double *s, int n) { /* split_step(n+1-k) = split_step(k) */
    for (int k = 0; k < n / 2; ++k) s[n - 1 - k] = s[k];
}

/* steps: room for 32 doubles.  Returns 0, or -1 for an unknown case. */
int orc_splitting_coeff(int split, double dt, double *s, int *nb_split_step, int *split_begin_T, int *dim_split_V) {
    int nb = 0, beginT = 0, dimV = 1;
    const double dt2 = dt * dt, dt4 = dt2 * dt2, dt6 = dt4 * dt2;
    memset(s, 0, sizeof(double) * 32);
    switch (split) {
    case 2: nb = 2; beginT = 1; s[0] = 1.0; s[1] = 1.0; break;                       /* sll_p_lie_tv :169-174 */
    case 3: nb = 2; beginT = 0; s[0] = 1.0; s[1] = 1.0; break;                       /* sll_p_lie_vt :175-180 */
    case 1: nb = 3; beginT = 1; s[0] = 0.5; s[1] = 1.0; s[2] = s[0]; break;          /* sll_p_strang_tvt :181-187 */
    case 0: nb = 3; beginT = 0; s[0] = 0.5; s[1] = 1.0; s[2] = s[0]; break;          /* sll_p_strang_vtv :188-194 */
    case 4: case 5:                                                                  /* triple jump :195-216 */
        nb = 7; beginT = (split == 4);
        s[0] = 0.675603595979829; s[1] = 1.351207191959658; s[2] = -0.17560359597982855; s[3] = -1.702414383919315;
        mirror(s, 7);
        break;
    case 6:                                                                          /* sll_p_order6_vtv :217-243 */
        nb = 23; beginT = 0;
        s[0] = 0.0414649985182624; s[1] = 0.123229775946271; s[2] = 0.198128671918067; s[3] = 0.290553797799558;
        s[4] = -0.0400061921041533; s[5] = -0.127049212625417; s[6] = 0.0752539843015807; s[7] = -0.246331761062075;
        s[8] = -0.0115113874206879; s[9] = 0.357208872795928; s[10] = 0.23666992478693111; s[11] = 0.20477705429147008;
        mirror(s, 23);
        break;
    case 7:                                                                          /* sll_p_order6_tvt :244-276 */
        nb = 29; beginT = 1;
        s[0] = 0.0378593198406116; s[1] = 0.09171915262446165; s[2] = 0.102635633102435; s[3] = 0.183983170005006;
        s[4] = -0.0258678882665587; s[5] = -0.05653436583288827; s[6] = 0.314241403071447; s[7] = 0.004914688774712854;
        s[8] = -0.130144459517415; s[9] = 0.143761127168358; s[10] = 0.106417700369543; s[11] = 0.328567693746804;
        s[12] = -0.00879424312851058; s[13] = -0.196411466486454234; s[14] = 0.20730506905689536;
        mirror(s, 29);
        break;
    case 8:                                                                          /* sll_p_order6vp_tvt :278-299 */
        nb = 9; beginT = 1;
        s[0] = 0.1095115577513980413559540;
        s[1] = 0.268722208204814693684441 - 2. * dt2 * 0.000805681667096178271312 + 4. * dt4 * 0.000017695766224036466792;
        s[2] = 0.4451715080955340951457244;
        s[3] = 0.2312777917951853063155588 - 2. * dt2 * 0.003955911930042478239763 + 4. * dt4 * 0.000052384078562246674986;
        s[4] = -0.1093661316938642730033570;
        mirror(s, 9);
        break;
    case 9:                                                                          /* sll_p_order6vp_vtv :300-323 */
        nb = 9; beginT = 0;
        s[0] = 0.359950808794143627485664 - 2. * dt2 * (-0.01359558332625151635) + 4. * dt4 * (-8.562814848565929e-6);
        s[1] = 1.079852426382430882456991;
        s[2] = -0.1437147273026540434771131 - 2. * dt2 * (-0.00385637757897273261) + 4. * dt4 * (0.0004883788785819335822);
        s[3] = -0.579852426382430882456991;
        s[4] = 0.567527837017020831982899 - 2. * dt2 * (-0.03227361602037480885) + 4. * dt4 * 0.002005141087312622342;
        mirror(s, 9);
        break;
    case 10:                                                                         /* sll_p_order6vpnew_tvt :324-347 */
        nb = 9; beginT = 1;
        s[0] = 0.1095115577513980413559540;
        s[1] = 0.268722208204814693684441 - 2. * dt2 * 0.000805681667096178271312 + 4. * dt4 * (8.643923349886021963e-6)
               - 8. * dt6 * (1.4231479258353431522e-6);
        s[2] = 0.4451715080955340951457244;
        s[3] = 0.2312777917951853063155588 - 2. * dt2 * 0.003955911930042478239763 + 4. * dt4 * (0.000061435921436397119815);
        s[4] = -0.1093661316938642730033570;
        mirror(s, 9);
        break;
    case 11:                                                                         /* sll_p_order6vpnew1_vtv :348-373 */
        nb = 11; beginT = 0;
        s[0] = 0.0490864609761162454914412 - 2. * dt2 * (0.0000697287150553050840999);
        s[1] = 0.1687359505634374224481957;
        s[2] = 0.2641776098889767002001462 - 2. * dt2 * (0.000625704827430047189169) + 4. * dt4 * (-2.91660045768984781644e-6);
        s[3] = 0.377851589220928303880766;
        s[4] = 0.1867359291349070543084126 - 2. * dt2 * (0.00221308512404532556163) + 4. * dt4 * (0.0000304848026170003878868)
               - 8. * dt6 * (4.98554938787506812159e-7);
        s[5] = -0.0931750795687314526579244;
        mirror(s, 11);
        break;
    case 12:                                                                         /* sll_p_order6vpnew2_vtv :563-589 */
        nb = 11; beginT = 0;
        s[0] = 0.083335463273305120964507 - 2. * dt2 * (-0.00015280483587048489661) + 4. * dt4 * (-0.0017675734111895638156)
               - 8. * dt6 * (0.00021214072262165668039);
        s[1] = 0.72431592569108212422250;
        s[2] = 0.827694857845135145869413 - 2. * dt2 * (-0.010726848627286273332) + 4. * dt4 * (0.012324362982853212700);
        s[3] = -0.4493507217041624582458844;
        s[4] = -0.4110303211184402668339201 - 2. * dt2 * (0.014962337009932798678);
        s[5] = 0.4500695920261606680467717;
        mirror(s, 11);
        break;
    case 13:                                                                         /* sll_p_order6vp2d_vtv :374-396 */
        nb = 11; beginT = 0;
        s[0] = 0.0490864609761162454914412 + 2. * dt2 * (0.00166171386175851683711044);
        s[1] = 0.1687359505634374224481957;
        s[2] = 0.2641776098889767002001462 - 2. * dt2 * (0.00461492847770001641230401);
        s[3] = 0.377851589220928303880766;
        s[4] = 0.1867359291349070543084126 + 2. * dt2 * (0.0000446959494108217402966857);
        s[5] = -0.0931750795687314526579244;
        mirror(s, 11);
        break;
    case 14:                                                                         /* sll_p_order6vpot_vtv :397-426 */
        nb = 11; beginT = 0; dimV = 2;
        s[0] = 0.0490864609761162454914412 + 2. * dt2 * (0.00166171386175851683711044);
        s[1] = dt2 * (0.00166171386175851683711044);
        s[2] = 0.1687359505634374224481957;
        s[3] = 0.2641776098889767002001462 - 2. * dt2 * (0.00461492847770001641230401);
        s[4] = -dt2 * (0.00461492847770001641230401);
        s[5] = 0.377851589220928303880766;
        s[6] = 0.1867359291349070543084126 + 2. * dt2 * (0.0000446959494108217402966857);
        s[7] = dt2 * (0.0000446959494108217402966857);
        s[8] = -0.0931750795687314526579244;
        s[9] = s[6]; s[10] = s[7]; s[11] = s[5]; s[12] = s[3]; s[13] = s[4]; s[14] = s[2]; s[15] = s[0]; s[16] = s[1];
        break;
    case 15:                                                                         /* sll_p_order6vpotnew1_vtv :428-466 */
        nb = 9; beginT = 0; dimV = 2;
        s[0] = 0.359950808794143627485664 + 2. * dt2 * (0.);
        s[1] = dt2 * (0.);
        s[2] = 1.079852426382430882456991;
        s[3] = -0.1437147273026540434771131 + 2. * dt2 * (0.0139652542242388403673);
        s[4] = dt2 * (0.0139652542242388403673);
        s[5] = -0.579852426382430882456991;
        s[6] = 0.567527837017020831982899 + 2. * dt2 * (0.039247029382345626020);
        s[7] = dt2 * (0.039247029382345626020);
        s[8] = s[5]; s[9] = s[3]; s[10] = s[4]; s[11] = s[2]; s[12] = s[0]; s[13] = s[1];
        break;
    case 16:                                                                         /* sll_p_order6vpotnew2_vtv :468-511 */
        nb = 11; beginT = 0; dimV = 2;
        s[0] = 0.086971698963920047813358 + 2. * dt2 * (1.98364114652831655458915e-6);
        s[1] = dt2 * (1.98364114652831655458915e-6);
        s[2] = 0.303629319055488881944104;
        s[3] = 0.560744966588102145251453 - 2. * dt2 * (0.00553752115152236516667268);
        s[4] = -dt2 * (0.00553752115152236516667268);
        s[5] = 0.303629319055488881944104;
        s[6] = -0.1477166655520221930648117 - 2. * dt2 * (0.00284218110811634663914191);
        s[7] = -dt2 * (0.00284218110811634663914191);
        s[8] = -0.2145172762219555277764167;
        s[9] = s[6]; s[10] = s[7]; s[11] = s[5]; s[12] = s[3]; s[13] = s[4]; s[14] = s[2]; s[15] = s[0]; s[16] = s[1];
        break;
    case 17:                                                                         /* sll_p_order6vpotnew3_vtv :513-561 */
        nb = 13; beginT = 0; dimV = 2;
        s[0] = 0.0482332301753032567427580 + 2. * dt2 * (0.0002566567904012107264);
        s[1] = dt2 * (0.0002566567904012107264);
        s[2] = 0.2701015188126056215752542;
        s[3] = 0.0482332301753032567427580 + 2. * dt2 * (0.0009439771580927593579);
        s[4] = dt2 * (0.0009439771580927593579);
        s[5] = -0.108612186368692920020654;
        s[6] = 0.2361392603742494444753990 - 2. * dt2 * (0.002494619878121813220);
        s[7] = -dt2 * (0.002494619878121813220);
        s[8] = 0.3385106675560872984454001;
        s[9] = 0.3347885585502880840781703 - 2. * dt2 * (0.002670269183371982607658111);
        s[10] = -dt2 * (0.002670269183371982607658111);
        s[11] = s[8]; s[12] = s[6]; s[13] = s[7]; s[14] = s[5]; s[15] = s[3]; s[16] = s[4]; s[17] = s[2]; s[18] = s[0]; s[19] = s[1];
        break;
    default: return -1;
    }
    *nb_split_step = nb; *split_begin_T = beginT; *dim_split_V = dimV;
    return 0;
}

/* sll_s_compute_w_hermite (sll_m_fcisl.F90:413-486): first-derivative finite-difference weights on the stencil
 * r..s (r < 0 < s), w indexed w[k - r] */
void orc_compute_w_hermite(int r, int s, double *w) {
    for (int i = r; i <= s; ++i) {
        if (i == 0) continue;
        double tmp = 1.0;
        for (int j = r; j <= i - 1; ++j) tmp *= (double)(i - j);
        for (int j = i + 1; j <= s; ++j) tmp *= (double)(i - j);
        tmp = 1.0 / tmp;
        for (int j = r; j <= s; ++j) if (j != i && j != 0) tmp *= (double)(-j);
        w[i - r] = tmp;
    }
    double tmp = 0.0;
    for (int i = r; i <= -1; ++i) tmp += w[i - r];
    for (int i = 1; i <= s; ++i) tmp += w[i - r];
    w[-r] = -tmp;
}

/* compute_jacobian (...poisson_serial.F90:1403-1436): E arrays are (nc1+1) x (nc2+1) with duplicated end points */
void orc_compute_jacobian(const double *E1, const double *E2, int nc1, int nc2, double factor, int r, int s, double *jac) {
    double *w = (double *)malloc(sizeof(double) * (s - r + 1));
    orc_compute_w_hermite(r, s, w);
    const int np1 = nc1 + 1;
    for (int j = 1; j <= nc2 + 1; ++j)
        for (int i = 1; i <= nc1 + 1; ++i) {
            double g11 = 0, g12 = 0, g21 = 0, g22 = 0;
            for (int k = r; k <= s; ++k) {
                int ii = ((i + k - 1 + nc1) % nc1 + nc1) % nc1 + 1, jj = ((j + k - 1 + nc2) % nc2 + nc2) % nc2 + 1;
                g11 += w[k - r] * E1[(ii - 1) + (long)np1 * (j - 1)];
                g12 += w[k - r] * E2[(ii - 1) + (long)np1 * (j - 1)];
                g21 += w[k - r] * E1[(i - 1) + (long)np1 * (jj - 1)];
                g22 += w[k - r] * E2[(i - 1) + (long)np1 * (jj - 1)];
            }
            jac[(i - 1) + (long)np1 * (j - 1)] = (g11 * g22 - g12 * g21) * factor;
        }
    free(w);
}
