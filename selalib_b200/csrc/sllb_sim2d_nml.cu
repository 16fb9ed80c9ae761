// sllb_sim2d_nml.cu -- namelist front-end and output files of the 1D1V simulation:
// the file sim_bsl_vp_1d1v_cart reads (simulations/parallel/bsl_vp_1d1v_cart/sll_m_sim_bsl_vp_1d1v_cart.F90:431-494 with
// the defaults of :496-565) drives sllb_sim2d_*, and the run leaves behind what the reference leaves behind:
//   thdiag.dat                      '(8g25.15)' + 3 (nb_mode+1) x '(1g25.15)' per diagnostic step (:1778-1801)
//   x.bdat, v.bdat                  node positions without the duplicated end point (:1261-1266)
//   f0.bdat                         the equilibrium (Landau initializer with eps = 0, :727-731,1346-1348)
//   deltaf.bdat                     f - f_equilibrium at t = 0 and every freq_diag steps (:1455-1459,1827-1828)
//   rhotot.bdat, efield.bdat, t.bdat  appended at t = 0 and at every diagnostic step (:1432-1434,1802-1804)
//   f_plot_<iplot>_proc_0000.rst    every freq_diag_restart steps: time + f with its duplicated end points (:1762-1770)
// and `restart_file` / `time_init_from_restart_file` are honoured on the way in (:1281-1312).
// .bdat/.rst are raw streams of doubles (sll_s_binary_write_array_*: ACCESS="STREAM", unformatted).
// Not offered (SLLB_ERR_UNSUPPORTED with the reference's wording): SLL_TWO_GRID_MESH, SLL_BEAM, the KEEN / Ampere drives,
// SLL_VLASOV_AMPERE, SLL_CONSERVATIVE advection form, the polar Poisson solver, the *VP* splittings.
// Host code only, on top of the public C ABI.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sllb_internal.h"
#include "sllb_namelist.h"

using namespace sllb;
using namespace sllb::namelist;

namespace {
struct Sim2dNml {
    int nit = 600, freq_diag = 100, freq_diag_time = 1, freq_diag_restart = 5000, nb_mode = 5;
    double time_init = 0.0;
    bool time_init_from_restart_file = false;
    std::string restart_file = "no_restart_file";
    double kmode = 0.5;
};
bool write_raw(const std::string &path, const char *mode, const double *d, size_t n) {
    FILE *fp = fopen(path.c_str(), mode);
    if (!fp) return false;
    const bool ok = fwrite(d, sizeof(double), n, fp) == n;
    fclose(fp);
    return ok;
}
int create_from_namelist(const char *filename, sllb_sim2d_t *S, Sim2dNml *out) {
    Namelist nml;
    std::string err, path(filename);
    FILE *probe = fopen(path.c_str(), "r");
    if (probe) fclose(probe);
    else path += ".nml"; // the reference appends the extension (:596)
    if (!parse_namelist(path.c_str(), nml, err)) return fail(SLLB_ERR_INVALID, "#init_vp2d_par_cart() " + err);
    const double pi = 3.14159265358979323846;
    // &initial_function (:523-529)
    const std::string ifc = get_str(nml, "initial_function", "initial_function_case", "SLL_LANDAU");
    int init;
    if (ifc == "SLL_LANDAU") init = 0;
    else if (ifc == "SLL_TWO_STREAM_INSTABILITY") init = 1;
    else if (ifc == "SLL_BUMP_ON_TAIL") init = 2;
    else return fail(SLLB_ERR_UNSUPPORTED, "#init_func_case not implemented: " + ifc);
    const double kmode = get_real(nml, "initial_function", "kmode", 0.5), eps = get_real(nml, "initial_function", "eps", 0.001);
    out->kmode = kmode;
    out->restart_file = get_str(nml, "initial_function", "restart_file", "no_restart_file");
    out->time_init_from_restart_file = get_bool(nml, "initial_function", "time_init_from_restart_file", false);
    // &geometry (:498-521,617-690); the declared default of mesh_case_x2 ends up as SLL_TWO_GRID_MESH (:509)
    const std::string m1 = get_str(nml, "geometry", "mesh_case_x1", "SLL_LANDAU_MESH");
    const std::string m2 = get_str(nml, "geometry", "mesh_case_x2", "SLL_TWO_GRID_MESH");
    const int n1 = get_int(nml, "geometry", "num_cells_x1", 32), n2 = get_int(nml, "geometry", "num_cells_x2", 64);
    const double x1min = get_real(nml, "geometry", "x1_min", 0.0);
    double x1max;
    if (m1 == "SLL_LANDAU_MESH") x1max = (double)get_int(nml, "geometry", "nbox_x1", 1) * 2.0 * pi / kmode;
    else if (m1 == "SLL_CARTESIAN_MESH") {
        if (!find(nml, "geometry", "x1_max")) return fail(SLLB_ERR_INVALID, "#x1_max must be given with mesh_case_x1 = SLL_CARTESIAN_MESH");
        x1max = get_real(nml, "geometry", "x1_max", 0.0);
    } else return fail(SLLB_ERR_UNSUPPORTED, "#mesh_case_x1 " + m1 + " not implemented");
    if (m2 != "SLL_CARTESIAN_MESH") return fail(SLLB_ERR_UNSUPPORTED, "#mesh_case_x2 " + m2 + " not implemented");
    if (get_int(nml, "geometry", "every_x1", 1) != 1 || get_int(nml, "geometry", "every_x2", 1) != 1)
        return fail(SLLB_ERR_UNSUPPORTED, "#every_x1 / every_x2 other than 1 not implemented");
    const double x2min = get_real(nml, "geometry", "x2_min", -6.0), x2max = get_real(nml, "geometry", "x2_max", 6.0);
    // &time_iterations (:531-539)
    const double dt = get_real(nml, "time_iterations", "dt", 0.1);
    out->nit = get_int(nml, "time_iterations", "number_iterations", 600);
    out->freq_diag = get_int(nml, "time_iterations", "freq_diag", 100);
    out->freq_diag_time = get_int(nml, "time_iterations", "freq_diag_time", 1);
    out->freq_diag_restart = get_int(nml, "time_iterations", "freq_diag_restart", 5000);
    out->nb_mode = get_int(nml, "time_iterations", "nb_mode", 5);
    out->time_init = get_real(nml, "time_iterations", "time_init", 0.0);
    if (out->nb_mode < 0) return fail(SLLB_ERR_INVALID, "#bad value of nb_mode; #should be >=0");
    if (out->freq_diag < 1 || out->freq_diag_time < 1 || out->freq_diag_restart < 1) return fail(SLLB_ERR_INVALID, "#freq_diag* must be >= 1");
    int split = 0;
    std::string sc = get_str(nml, "time_iterations", "split_case", "SLL_STRANG_VTV");
    // this simulation spells them SLL_ORDER6VPNEW_TVT / SLL_ORDER6VPNEW1_VTV / SLL_ORDER6VPNEW2_VTV (:840-845), the 2D2V one
    // (whose table the library holds) SLL_ORDER6VPnew...
    const size_t pos = sc.find("VPNEW");
    if (pos != std::string::npos) sc.replace(pos, 5, "VPnew");
    if (sllb_splitting_case_from_name(sc.c_str(), &split)) return fail(SLLB_ERR_INVALID, "#split_case not defined");
    // &advector (:548-556,866-925)
    int mm[2], oo[2];
    const char *av[2] = {"advector_x1", "advector_x2"}, *od[2] = {"order_x1", "order_x2"};
    for (int d = 0; d < 2; ++d) {
        const std::string a = get_str(nml, "advector", av[d], "SLL_LAGRANGE");
        oo[d] = get_int(nml, "advector", od[d], 4);
        if (a == "SLL_SPLINES") {
            if (oo[d] != 4 && oo[d] != 6 && oo[d] != 8) return fail(SLLB_ERR_UNSUPPORTED, std::string("#") + av[d] + ": periodic splines are implemented for orders 4, 6 and 8");
            mm[d] = SLLB_METHOD_SPLINE;
        } else if (a == "SLL_LAGRANGE") {
            if (oo[d] < 4 || oo[d] > 18 || oo[d] % 2 != 0) return fail(SLLB_ERR_UNSUPPORTED, std::string("#") + av[d] + ": periodic Lagrange is implemented for even orders 4 .. 18");
            mm[d] = SLLB_METHOD_LAGRANGE_CENTERED;
        } else return fail(SLLB_ERR_UNSUPPORTED, std::string("#") + av[d] + " " + a + " not implemented");
    }
    if (get_str(nml, "advector", "advection_form_x2", "SLL_ADVECTIVE") != "SLL_ADVECTIVE")
        return fail(SLLB_ERR_UNSUPPORTED, "#advection_form_x2 SLL_CONSERVATIVE not implemented");
    if (get_str(nml, "advector", "integration_case", "SLL_TRAPEZOID") != "SLL_TRAPEZOID")
        return fail(SLLB_ERR_UNSUPPORTED, "#integration_case: SLL_TRAPEZOID only");
    if (get_real(nml, "advector", "factor_x1", 1.0) != 1.0 || get_real(nml, "advector", "factor_x2_rho", 1.0) != 1.0 ||
        get_real(nml, "advector", "factor_x2_1", 1.0) != 1.0)
        return fail(SLLB_ERR_UNSUPPORTED, "#factor_x1 / factor_x2_rho / factor_x2_1 other than 1 not implemented");
    if (get_str(nml, "poisson", "poisson_solver", "SLL_FFT") != "SLL_FFT") return fail(SLLB_ERR_UNSUPPORTED, "#poisson_solver: SLL_FFT only");
    if (get_str(nml, "drive", "drive_type", "SLL_NO_DRIVE") != "SLL_NO_DRIVE") return fail(SLLB_ERR_UNSUPPORTED, "#drive_type: SLL_NO_DRIVE only");
    SLLB_TRY(sllb_sim2d_create(n1, n2, x1min, x1max, x2min, x2max, init, kmode, eps, dt, mm[0], oo[0], S));
    int rc = sllb_sim2d_set_advectors(*S, mm[0], oo[0], mm[1], oo[1]);
    if (!rc) rc = sllb_sim2d_set_splitting(*S, split);
    if (rc) { sllb_sim2d_destroy(*S); *S = nullptr; return rc; }
    return SLLB_OK;
}
} // namespace

extern "C" {

int sllb_sim2d_create_from_namelist(const char *filename, sllb_sim2d_t *S, int *number_iterations, int *freq_diag_time,
                                    int *nb_mode) {
    if (!filename || !S) return fail(SLLB_ERR_INVALID, "sim2d_create_from_namelist: null");
    SLLB_TRY(require_device());
    Sim2dNml n;
    SLLB_TRY(create_from_namelist(filename, S, &n));
    if (number_iterations) *number_iterations = n.nit;
    if (freq_diag_time) *freq_diag_time = n.freq_diag_time;
    if (nb_mode) *nb_mode = n.nb_mode;
    return SLLB_OK;
}

/* the whole program sim_bsl_vp_1d1v_cart (single process): namelist in, the reference's files out (into `outdir`) */
int sllb_sim2d_run_namelist(const char *filename, const char *outdir) {
    if (!filename) return fail(SLLB_ERR_INVALID, "sim2d_run_namelist: null");
    SLLB_TRY(require_device());
    const std::string dir = (outdir && outdir[0]) ? std::string(outdir) + "/" : std::string();
    sllb_sim2d_t S = nullptr;
    Sim2dNml n;
    SLLB_TRY(create_from_namelist(filename, &S, &n));
    int rc = SLLB_OK;
    sllb_field_t F = nullptr;
    rc = sllb_sim2d_field(S, &F);
    int ext[6] = {0};
    int ndim = 0;
    if (!rc) rc = sllb_field_extents(F, &ndim, ext);
    const int n1 = ext[0], n2 = ext[1];
    std::vector<double> buf, feq;
    double xlim[4];
    if (!rc) rc = sllb_sim2d_geometry(S, xlim);
    if (!rc) {
        // x.bdat / v.bdat (:1261-1266)
        std::vector<double> x(n1), v(n2);
        for (int i = 0; i < n1; ++i) x[i] = xlim[0] + i * (xlim[1] - xlim[0]) / n1;
        for (int j = 0; j < n2; ++j) v[j] = xlim[2] + j * (xlim[3] - xlim[2]) / n2;
        if (!write_raw(dir + "x.bdat", "wb", x.data(), x.size()) || !write_raw(dir + "v.bdat", "wb", v.data(), v.size()))
            rc = fail(SLLB_ERR_INVALID, "sim2d_run_namelist: cannot write x.bdat / v.bdat");
        // equilibrium = sll_f_landau_initializer_2d with eps = 0 (:727-731): f0.bdat
        feq.resize((size_t)n1 * n2);
        const double fac = 1.0 / sqrt(2.0 * 3.14159265358979323846);
        for (int j = 0; j < n2; ++j)
            for (int i = 0; i < n1; ++i) feq[i + (size_t)n1 * j] = fac * (1.0 + 0.0 * cos(n.kmode * x[i])) * exp(-0.5 * v[j] * v[j]);
        if (!rc && !write_raw(dir + "f0.bdat", "wb", feq.data(), feq.size())) rc = fail(SLLB_ERR_INVALID, "sim2d_run_namelist: cannot write f0.bdat");
    }
    // restart file (:1281-1312)
    double time_init = n.time_init;
    if (!rc && n.restart_file != "no_restart_file") {
        double t = 0.0;
        rc = sllb_sim2d_read_restart(S, (n.restart_file + "_proc_0000.rst").c_str(), &t);
        if (!rc && n.time_init_from_restart_file) time_init = t;
    }
    if (!rc) rc = sllb_sim2d_set_time(S, time_init);
    FILE *th = nullptr;
    if (!rc) {
        th = fopen((dir + "thdiag.dat").c_str(), "w");
        if (!th) rc = fail(SLLB_ERR_INVALID, "sim2d_run_namelist: cannot create thdiag.dat");
    }
    const int ncol = 8 + 3 * (n.nb_mode + 1);
    std::vector<double> row(ncol), hr(n1), hE(n1);
    buf.resize((size_t)n1 * n2);
    auto append_fields = [&](double t, const char *mode) -> int {
        SLLB_TRY(sllb_sim2d_fields_host(S, hr.data(), hE.data()));
        if (!write_raw(dir + "efield.bdat", mode, hE.data(), hE.size()) || !write_raw(dir + "rhotot.bdat", mode, hr.data(), hr.size()) ||
            !write_raw(dir + "t.bdat", mode, &t, 1))
            return fail(SLLB_ERR_INVALID, "sim2d_run_namelist: cannot write efield.bdat / rhotot.bdat / t.bdat");
        return SLLB_OK;
    };
    auto append_deltaf = [&](const char *mode) -> int {
        SLLB_TRY(sllb_field_download(F, buf.data(), nullptr));
        for (size_t k = 0; k < buf.size(); ++k) buf[k] -= feq[k];
        if (!write_raw(dir + "deltaf.bdat", mode, buf.data(), buf.size())) return fail(SLLB_ERR_INVALID, "sim2d_run_namelist: cannot write deltaf.bdat");
        return SLLB_OK;
    };
    // t = 0: fields and deltaf; the first t.bdat entry is istep*dt = 0 whatever time_init is (:1434)
    if (!rc) rc = append_fields(0.0, "wb");
    if (!rc) rc = append_deltaf("wb");
    int iplot = 1;   // sll_v iplot starts at 1 (:1170)
    for (int it = 1; it <= n.nit && !rc; ++it) {
        rc = sllb_sim2d_run(S, 1, nullptr);
        if (rc || it % n.freq_diag_time != 0) continue;
        rc = sllb_sim2d_thdiag(S, n.nb_mode, row.data());
        if (rc) break;
        if (it % n.freq_diag_restart == 0) {
            char name[64];
            snprintf(name, sizeof(name), "f_plot_%04d_proc_0000.rst", iplot);
            rc = sllb_sim2d_write_restart(S, (dir + name).c_str());
            if (rc) break;
        }
        char cell[64];
        for (int k = 0; k < ncol && !rc; ++k) {
            rc = sllb_format_g(row[k], 25, 15, cell);
            if (!rc) fputs(cell, th);
        }
        fputc('\n', th);
        if (!rc) rc = append_fields(row[0], "ab");
        if (!rc && it % n.freq_diag == 0) {
            rc = append_deltaf("ab");
            iplot += 1;
        }
    }
    if (th) fclose(th);
    sllb_sim2d_destroy(S);
    return rc;
}

} // extern "C"
