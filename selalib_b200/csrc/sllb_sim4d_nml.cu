// sllb_sim4d_nml.cu -- namelist front-end and thdiag writer of the 2D2V simulation:
// the file sim_bsl_vp_2d2v_cart_poisson_serial reads (simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/
// sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:236-440: namelists geometry, initial_function, time_iterations,
// advector, poisson with their defaults; mesh cases :393-436; split_case :514-554; advectors :556-624) drives
// sllb_sim4d_*, and the time-history file is written like the reference's thdiag.dat ('(13g20.12)' rows, :998-1010,
// 1262-1275).  Host code only, on top of the public C ABI.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "sllb_internal.h"
#include "sllb_namelist.h"

using namespace sllb;
using namespace sllb::namelist;

extern "C" {

/* Fortran G20.12: F(16).(12-s) followed by four blanks when the value rounded to 12 significant digits lies in
 * [0.1, 10^12), E20.12 otherwise, zero as F16.11 (Fortran 2003, 10.6.4.1.2) */
int sllb_format_g20_12(double x, char *buf21) { return sllb_format_g(x, 20, 12, buf21); }
/* Fortran Gw.d (no exponent width): used as g20.12 by the 2D2V thdiag file and as g25.15 by the 1D1V one */
int sllb_format_g(double x, int w, int d, char *buf) {
    if (!buf || w < d + 7 || w > 60 || d < 1) return fail(SLLB_ERR_INVALID, "format_g: need d >= 1 and d + 7 <= w <= 60");
    char body[160];
    if (x == 0.0) snprintf(body, sizeof(body), "%*.*f    ", w - 4, d - 1, 0.0);
    else if (!std::isfinite(x)) snprintf(body, sizeof(body), "%*s", w, std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity"));
    else {
        char e[64];
        snprintf(e, sizeof(e), "%.*e", d - 1, fabs(x)); // m.mmmmmmmmmmme+XX, rounded to 12 significant digits
        const char *ep = strchr(e, 'e');
        const int s = atoi(ep + 1) + 1;                  // rounded value = 0.mmm... * 10^s
        if (s >= 0 && s <= d) {
            char f[64];
            snprintf(f, sizeof(f), d - s == 0 ? "%.*f." : "%.*f", d - s, x); // F16.0 keeps the decimal point
            snprintf(body, sizeof(body), "%*s    ", w - 4, f);
        } else {
            char digits[64];   // d <= 53 significant digits (w <= 60)
            int nd = 0;
            for (const char *c = e; c < ep; ++c) if (*c >= '0' && *c <= '9') digits[nd++] = *c;
            digits[nd] = 0;
            char m[64];
            snprintf(m, sizeof(m), "%s0.%sE%c%02d", x < 0 ? "-" : "", digits, s < 0 ? '-' : '+', abs(s));
            snprintf(body, sizeof(body), "%*s", w, m);
        }
    }
    if ((int)strlen(body) != w) { // does not fit: Fortran prints asterisks
        memset(body, '*', w); body[w] = 0;
    }
    memcpy(buf, body, w + 1);
    return SLLB_OK;
}

int sllb_sim4d_create_from_namelist(const char *filename, sllb_comm_t comm, sllb_sim4d_t *S, int *number_iterations,
                                    int *freq_diag_time) {
    if (!filename || !S) return fail(SLLB_ERR_INVALID, "sim4d_create_from_namelist: null");
    Namelist nml;
    std::string err, path(filename);
    FILE *probe = fopen(path.c_str(), "r");
    if (probe) fclose(probe);
    else path += ".nml"; // the reference appends the extension (:375)
    if (!parse_namelist(path.c_str(), nml, err))
        return fail(SLLB_ERR_INVALID, "#initialize_vlasov_par_poisson_seq_cart() " + err);
    sllb_sim4d_params_t p;
    memset(&p, 0, sizeof(p));
    const double pi = 3.14159265358979323846;
    // &initial_function (:329-332)
    const std::string ifc = get_str(nml, "initial_function", "initial_function_case", "SLL_LANDAU");
    if (ifc != "SLL_LANDAU") return fail(SLLB_ERR_UNSUPPORTED, "#init_func_case not implemented: " + ifc + " (SLL_LANDAU only)");
    p.kx1 = get_real(nml, "initial_function", "kmode_x1", 0.5);
    p.kx2 = get_real(nml, "initial_function", "kmode_x2", 0.5);
    p.eps = get_real(nml, "initial_function", "eps", 1e-3);
    // &geometry (:311-327,393-436)
    const char *mc[4] = {"mesh_case_x1", "mesh_case_x2", "mesh_case_x3", "mesh_case_x4"};
    const char *nc[4] = {"num_cells_x1", "num_cells_x2", "num_cells_x3", "num_cells_x4"};
    const char *mn[4] = {"x1_min", "x2_min", "x3_min", "x4_min"}, *mx[4] = {"x1_max", "x2_max", "x3_max", "x4_max"};
    const char *nb[2] = {"nbox_x1", "nbox_x2"};
    for (int d = 0; d < 4; ++d) {
        const std::string mesh = get_str(nml, "geometry", mc[d], d < 2 ? "SLL_LANDAU_MESH" : "SLL_CARTESIAN_MESH");
        p.nc[d] = get_int(nml, "geometry", nc[d], 16);
        p.xmin[d] = get_real(nml, "geometry", mn[d], d < 2 ? 0.0 : -6.0);
        if (mesh == "SLL_LANDAU_MESH" && d < 2)
            p.xmax[d] = (double)get_int(nml, "geometry", nb[d], 1) * 2.0 * pi / (d == 0 ? p.kx1 : p.kx2);
        else if (mesh == "SLL_CARTESIAN_MESH") {
            // x3_max / x4_max default to 6 (:323,327); x1_max / x2_max have no default in the reference (:314-321: the
            // variable is read before it is ever set), so a namelist that leaves them out is refused here
            if (d < 2 && !find(nml, "geometry", mx[d]))
                return fail(SLLB_ERR_INVALID, std::string("#") + mx[d] + " must be given with " + mc[d] + " = SLL_CARTESIAN_MESH "
                                              "(the reference has no default for it)");
            p.xmax[d] = get_real(nml, "geometry", mx[d], 6.0);
        }
        else
            return fail(SLLB_ERR_UNSUPPORTED, std::string("#") + mc[d] + " " + mesh + " not implemented");
    }
    // &time_iterations (:334-338)
    p.dt = get_real(nml, "time_iterations", "dt", 2.0);
    const int nit = get_int(nml, "time_iterations", "number_iterations", 5);
    const int fdt = get_int(nml, "time_iterations", "freq_diag_time", 1);
    const std::string sc = get_str(nml, "time_iterations", "split_case", "SLL_ORDER6VPnew1_VTV");
    SLLB_TRY(sllb_splitting_case_from_name(sc.c_str(), &p.split));
    // &advector (:349-356,556-624)
    const char *av[4] = {"advector_x1", "advector_x2", "advector_x3", "advector_x4"};
    const char *od[4] = {"order_x1", "order_x2", "order_x3", "order_x4"};
    for (int d = 0; d < 4; ++d) {
        const std::string a = get_str(nml, "advector", av[d], "SLL_LAGRANGE");
        const int order = get_int(nml, "advector", od[d], 4);
        if (a == "SLL_SPLINES") {
            if (order != 4 && order != 6 && order != 8) return fail(SLLB_ERR_UNSUPPORTED, std::string("#advector ") + av[d] + ": periodic splines are implemented for orders 4, 6 and 8");
            p.method_axis[d] = SLLB_METHOD_SPLINE;
        } else if (a == "SLL_LAGRANGE") {
            if (order < 4 || order > 18 || order % 2 != 0) return fail(SLLB_ERR_UNSUPPORTED, std::string("#advector ") + av[d] + ": periodic Lagrange is implemented for even orders 4 .. 18");
            p.method_axis[d] = SLLB_METHOD_LAGRANGE_CENTERED;
        } else
            return fail(SLLB_ERR_UNSUPPORTED, std::string("#advector in x") + char('1' + d) + " " + a + " not implemented");
        p.order_axis[d] = order;
    }
    p.method = p.method_axis[0]; p.order = p.order_axis[0];
    // &poisson (:358-359)
    p.stencil_r = get_int(nml, "poisson", "stencil_r", -2);
    p.stencil_s = get_int(nml, "poisson", "stencil_s", 2);
    SLLB_TRY(sllb_sim4d_create(&p, comm, S));
    if (number_iterations) *number_iterations = nit;
    if (freq_diag_time) *freq_diag_time = fdt > 0 ? fdt : 1;
    return SLLB_OK;
}

static int write_row13(FILE *fp, const double *row) {
    char buf[24];
    for (int k = 0; k < 13; ++k) {
        SLLB_TRY(sllb_format_g20_12(row[k], buf));
        fputs(buf, fp);
    }
    fputc('\n', fp);
    return SLLB_OK;
}

int sllb_sim4d_run_namelist(const char *filename, sllb_comm_t comm, const char *thdiag_path) {
    sllb_sim4d_t S = nullptr;
    int nit = 0, fdt = 1;
    SLLB_TRY(require_device());
    // Only rank 0 touches the file, but every rank takes part in the collectives of the time loop: the outcome of the
    // rank-0-only fopen is shared BEFORE anything collective starts, so that all ranks give up together; a failed write
    // later on is remembered and reported at the end without leaving the loop (no rank is left alone in a collective).
    const bool writer = !comm || comm->rank == 0;
    FILE *fp = nullptr;
    int open_failed = 0;
    if (writer) {
        fp = fopen(thdiag_path ? thdiag_path : "thdiag.dat", "w");
        open_failed = fp ? 0 : 1;
    }
    if (comm && comm->nranks > 1) {
        DevBuf flag;
        SLLB_TRY(flag.ensure(1));
        const double mine = (double)open_failed;
        double all = 0.0;
        SLLB_CUDA(cudaMemcpy(flag.p, &mine, sizeof(double), cudaMemcpyHostToDevice));
        SLLB_TRY(sllb_comm_allreduce_sum(comm, flag.p, 1));
        SLLB_CUDA(cudaMemcpy(&all, flag.p, sizeof(double), cudaMemcpyDeviceToHost));
        open_failed = all > 0.0 ? 1 : 0;
    }
    if (open_failed) { if (fp) fclose(fp); return fail(SLLB_ERR_INVALID, "sim4d_run_namelist: cannot create the thdiag file"); }
    int rc = sllb_sim4d_create_from_namelist(filename, comm, &S, &nit, &fdt);   // same namelist, same verdict on every rank
    if (rc) { if (fp) fclose(fp); return rc; }
    double row[13];
    int wrc = SLLB_OK;
    rc = sllb_sim4d_thdiag(S, row);
    if (!rc && fp) wrc = write_row13(fp, row);
    for (int it = 1; it <= nit && !rc; ++it) {
        rc = sllb_sim4d_run(S, 1, 0, nullptr);
        if (!rc && it % fdt == 0) {
            rc = sllb_sim4d_thdiag(S, row);
            if (!rc && fp && !wrc) wrc = write_row13(fp, row);
        }
    }
    if (fp) fclose(fp);
    sllb_sim4d_destroy(S);
    return rc ? rc : wrc;
}

} // extern "C"
