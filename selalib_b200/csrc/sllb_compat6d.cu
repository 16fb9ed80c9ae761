// sllb_compat6d.cu -- the C interface the reference's 6D simulation already exports
// (simulations/parallel/bsl_vp_3d3v_cart_dd/sll_m_sim_bsl_vp_3d3v_cart_dd_slim_interface.F90:63-283, driven by
// test_cpp_interface.cpp), re-exported with the same symbol names and by-reference signatures on top of
// sllb_sim6d_*, so a host program written against the Fortran simulation links against libsllb200.so instead.
// Includes a reader for the namelist file the simulation takes (sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:322-358,
// landau_params: sll_m_distribution_function_initializer_6d.F90:555) and the <prefix>.dat writer / ctest check
// (sll_m_sim_6d_utilities.F90:632-633,648-687).  Host code only.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sll_b200_sim6d_compat.h"
#include "sllb_internal.h"
#include "sllb_namelist.h"

using namespace sllb;
using namespace sllb::namelist;

namespace {

// Fortran e20.12 edit descriptor: 0.dddddddddddde+xx, right-justified in 20 columns
std::string fortran_e20_12(double x) {
    char out[64];
    if (x == 0.0) { snprintf(out, sizeof(out), "%20s", "0.000000000000E+00"); return out; }
    int ex = (int)floor(log10(fabs(x))) + 1;
    double m = x / pow(10.0, ex);
    char mant[32];
    snprintf(mant, sizeof(mant), "%.12f", fabs(m));
    if (mant[0] == '1') { // rounding carried into 1.000000000000
        ex += 1; m /= 10.0;
        snprintf(mant, sizeof(mant), "%.12f", fabs(m));
    }
    char body[48];
    snprintf(body, sizeof(body), "%s%sE%c%02d", x < 0 ? "-" : "", mant, ex < 0 ? '-' : '+', abs(ex));
    snprintf(out, sizeof(out), "%20s", body);
    return out;
}

struct Compat6d {
    sllb_sim6d *S = nullptr;
    sllb_sim6d_params_t p;
    int n_iterations = 0, first_time_step = 1, n_diagnostics = 1;
    bool ctest = false;
    std::string ctest_ref_file, prefix, nml_dir;
    FILE *dat = nullptr;
    int nw[6] = {0, 0, 0, 0, 0, 0}, mn[6] = {0, 0, 0, 0, 0, 0};
    std::vector<double> mirror;  // host mirror of the local block handed out by get_distribution
    bool mirror_out = false;     // the caller may have written into the mirror
    int rank = 0;
    bool clocks = false;         // SLLB_CLOCKS=1: the reference's stopwatch table, written to sll_clocks.txt by run
};

sllb_comm *g_compat_comm = nullptr;

[[noreturn]] void die(const char *fun, const std::string &msg) {
    // the reference's procedures return nothing: SLL_ERROR prints and stops the program
    fprintf(stderr, " ERROR in %s: %s\n", fun, msg.c_str());
    exit(1);
}
#define CK(call, fun)                                   \
    do {                                                \
        if ((call) != SLLB_OK) die(fun, sllb_last_error()); \
    } while (0)

Compat6d *self(void **sim, const char *fun) {
    if (!sim || !*sim) die(fun, "null simulation handle");
    return static_cast<Compat6d *>(*sim);
}
void push_mirror(Compat6d *c, const char *fun) {
    if (!c->mirror_out) return;
    sllb_field *F = nullptr;
    CK(sllb_sim6d_field(c->S, &F), fun);
    CK(sllb_field_upload(F, c->mirror.data(), nullptr), fun);
    c->mirror_out = false;
}
void write_row(Compat6d *c, const double *row14) {
    if (c->rank != 0 || !c->dat) return;
    std::string line;
    for (int k = 0; k < 14; ++k) line += fortran_e20_12(row14[k]);
    fprintf(c->dat, "%s\n", line.c_str());
    fflush(c->dat);
}

} // namespace

extern "C" {

/* multi-GPU: hand in the NCCL communicator before init (the reference hands in an MPI communicator through
 * sll_s_set_communicator_collective, sll_m_collective.F90:441-452) */
int sllb_sim6d_compat_set_comm(sllb_comm_t comm) {
    g_compat_comm = comm;
    return SLLB_OK;
}
/* MPI hand-over of the reference (sll_m_collective.F90:433-460): nothing to do without MPI */
void sll_s_allocate_collective(void) {}
void sll_s_set_communicator_collective(int *mpi_comm_f) { (void)mpi_comm_f; }
void sll_s_halt_collective(void) {}

void sim_bsl_vp_3d3v_cart_dd_slim_init(void **sim, const char *filename) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_init";
    if (!sim || !filename) die(fun, "null argument");
    Namelist nml;
    std::string err;
    if (!parse_namelist(filename, nml, err)) die(fun, "init_6d_vpB_dd_slim() " + err);
    Compat6d *c = new Compat6d();
    memset(&c->p, 0, sizeof(c->p));
    const double final_time = get_real(nml, "sim_params", "final_time", 0.0);
    c->p.delta_t = get_real(nml, "sim_params", "delta_t", 0.01);
    const int n_it = get_int(nml, "sim_params", "n_iterations", -1);
    // iff n_iterations is set it takes preference over final_time (:385-391)
    c->n_iterations = n_it < 0 ? (int)lround(final_time / c->p.delta_t) : n_it;
    c->ctest = get_bool(nml, "sim_params", "ctest", false);
    c->ctest_ref_file = get_str(nml, "sim_params", "ctest_ref_file", "");
    const std::string test_case = get_str(nml, "sim_params", "test_case", "landau_prod");
    if (test_case != "landau_prod") die(fun, "test_case '" + test_case + "' not implemented (landau_prod only)");
    const char *cells[6] = {"num_cells_x1", "num_cells_x2", "num_cells_x3", "num_cells_x4", "num_cells_x5", "num_cells_x6"};
    for (int d = 0; d < 6; ++d) c->p.n[d] = get_int(nml, "grid_dims", cells[d], 16);
    c->p.v_max = get_real(nml, "domain_dims", "v_max", 6.0);
    const char *xm[3] = {"x1_max", "x2_max", "x3_max"};
    for (int d = 0; d < 3; ++d) c->p.x_max[d] = get_real(nml, "domain_dims", xm[d], 12.5663706144);
    if (get_str(nml, "advect_params", "bc_type", "sll_p_periodic") != "sll_p_periodic") die(fun, "bc_type not implemented (sll_p_periodic only)");
    const std::string itype = get_str(nml, "advect_params", "interpolator_type", "fixed");
    if (itype == "fixed") c->p.advector = SLLB_ADVECTOR_FIXED;
    else if (itype == "centered") c->p.advector = SLLB_ADVECTOR_CENTERED;
    else if (itype == "spline" || itype == "splines") c->p.advector = SLLB_ADVECTOR_SPLINE;
    else die(fun, "Interpolator type not implemented.");
    c->p.stencil_v = get_int(nml, "advect_params", "stencil", 7);
    c->p.stencil_x = get_int(nml, "advect_params", "stencil_x", c->p.stencil_v);
    c->prefix = get_str(nml, "output", "file_prefix", "vp_3d3v_dd");
    c->p.time_in_phase = get_bool(nml, "output", "time_in_phase", true) ? 1 : 0;
    c->n_diagnostics = get_int(nml, "restart_params", "n_diagnostics", 1);
    int pg[6];
    for (int d = 0; d < 6; ++d) pg[d] = get_int(nml, "parallel_params", "process_grid", 0, d);
    c->p.alpha = get_real(nml, "landau_params", "alpha", 0.01);
    for (int d = 0; d < 3; ++d) {
        c->p.kx[d] = get_real(nml, "landau_params", "kx", 0.5, d);
        c->p.v_thermal[d] = get_real(nml, "landau_params", "v_thermal", 1.0, d);
    }
    std::string fn(filename);
    size_t slash = fn.find_last_of('/');
    c->nml_dir = slash == std::string::npos ? "." : fn.substr(0, slash);
    c->rank = g_compat_comm ? g_compat_comm->rank : 0;
    CK(sllb_sim6d_create_dist(&c->p, g_compat_comm, pg, &c->S), fun);
    {
        const char *e = getenv("SLLB_CLOCKS");
        c->clocks = e && e[0] == '1';
        if (c->clocks) CK(sllb_sim6d_set_clocks(c->S, 1), fun);
    }
    sllb_dd6d_t D = nullptr;
    CK(sllb_sim6d_decomposition(c->S, &D), fun);
    CK(sllb_dd6d_layout(D, nullptr, nullptr, c->mn, c->nw, nullptr, nullptr), fun);
    if (c->rank == 0) {
        printf(" Running 6D Vlasov simulation with domain decomposition (slim) on the B200 path ...\n");
        c->dat = fopen((c->prefix + ".dat").c_str(), "w");
        if (!c->dat) die(fun, "cannot create " + c->prefix + ".dat");
    }
    // diagnostics row at t = 0, written by init before any advection (:623-638)
    double row[14];
    CK(sllb_sim6d_diagnostics(c->S, 0.0, row), fun);
    write_row(c, row);
    *sim = c;
}

/* run_6d_vp_dd (:643-779): advect_v(dt/2), then n_iterations of {advect_x, rho + Poisson, diagnostics, advect_v} */
void sim_bsl_vp_3d3v_cart_dd_slim_run(void **sim) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_run";
    Compat6d *c = self(sim, fun);
    push_mirror(c, fun);
    CK(sllb_sim6d_advect_v(c->S, 0.5 * c->p.delta_t), fun);
    if (c->rank == 0) printf(" Entering main loop ... \n");
    const int last = c->first_time_step + c->n_iterations - 1;
    int itime;
    for (itime = c->first_time_step; itime <= last; ++itime) {
        CK(sllb_sim6d_advect_x(c->S), fun);
        CK(sllb_sim6d_prefetch_v_halo(c->S), fun);
        CK(sllb_sim6d_fields(c->S), fun);
        if (itime % c->n_diagnostics == 0) {
            double row[14];
            CK(sllb_sim6d_diagnostics(c->S, (double)itime * c->p.delta_t, row), fun);
            write_row(c, row);
        }
        if (c->p.time_in_phase && itime == last) CK(sllb_sim6d_advect_v(c->S, 0.5 * c->p.delta_t), fun);
        else CK(sllb_sim6d_advect_v(c->S, c->p.delta_t), fun);
        if (c->rank == 0) printf(" Time %7.3f of %7.3f\n", (double)itime * c->p.delta_t, (double)last * c->p.delta_t);
        FILE *stop = fopen("stop", "r"); // cooperative shutdown (sll_m_sim_6d_utilities.F90:732-750)
        if (stop) { fclose(stop); ++itime; break; }
    }
    c->first_time_step = itime;
    c->ctest = false; // like the interface's run (:101)
    CK(sllb_synchronize(), fun);
    // sll_s_finalize_clocks (:760): rank 0 leaves sll_clocks.txt behind when the stopwatches were on (SLLB_CLOCKS=1)
    if (c->rank == 0 && c->clocks) CK(sllb_sim6d_write_clocks(c->S, "sll_clocks.txt"), fun);
    if (c->rank == 0) printf(" Leaving main loop.\n");
}

void sim_bsl_vp_3d3v_cart_dd_slim_delete(void **sim) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_delete";
    Compat6d *c = self(sim, fun);
    if (c->dat) { fclose(c->dat); c->dat = nullptr; }
    if (c->rank == 0 && c->ctest) sllb_sim6d_compat_check((c->nml_dir + "/" + c->ctest_ref_file).c_str(), (c->prefix + ".dat").c_str());
    sllb_sim6d_destroy(c->S);
    delete c;
    *sim = nullptr;
}

/* the reference returns the LIVE array; here the live array is in HBM, so a host mirror of the local block is
 * refreshed and handed out, and pushed back to the device before the next compute call */
void sim_bsl_vp_3d3v_cart_dd_slim_get_distribution(void **sim, double **f) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_get_distribution";
    Compat6d *c = self(sim, fun);
    if (!f) die(fun, "null argument");
    push_mirror(c, fun);
    size_t total = 1;
    for (int d = 0; d < 6; ++d) total *= (size_t)c->nw[d];
    c->mirror.resize(total);
    sllb_field *F = nullptr;
    CK(sllb_sim6d_field(c->S, &F), fun);
    CK(sllb_field_download(F, c->mirror.data(), nullptr), fun);
    c->mirror_out = true;
    *f = c->mirror.data();
}
/* Fortran signature: type(c_ptr), VALUE (interface.F90:124-138) */
void sim_bsl_vp_3d3v_cart_dd_slim_set_distribution(void **sim, double *f) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_set_distribution";
    Compat6d *c = self(sim, fun);
    if (!f) die(fun, "null argument");
    sllb_field *F = nullptr;
    CK(sllb_sim6d_field(c->S, &F), fun);
    CK(sllb_field_upload(F, f, nullptr), fun);
    c->mirror_out = false;
}
void sim_bsl_vp_3d3v_cart_dd_slim_get_local_size(void **sim, int32_t *n6) {
    Compat6d *c = self(sim, "sim_bsl_vp_3d3v_cart_dd_slim_get_local_size");
    for (int d = 0; d < 6; ++d) n6[d] = c->nw[d];
}
void sim_bsl_vp_3d3v_cart_dd_slim_advect_v(void **sim, double *delta_t) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_advect_v";
    Compat6d *c = self(sim, fun);
    push_mirror(c, fun);
    CK(sllb_sim6d_advect_v(c->S, *delta_t), fun);
}
void sim_bsl_vp_3d3v_cart_dd_slim_advect_x(void **sim) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_advect_x";
    Compat6d *c = self(sim, fun);
    push_mirror(c, fun);
    CK(sllb_sim6d_advect_x(c->S), fun);
}
void sim_bsl_vp_3d3v_cart_dd_slim_print_etas(void **sim) {
    Compat6d *c = self(sim, "sim_bsl_vp_3d3v_cart_dd_slim_print_etas");
    const char *names[6] = {"x1", "x2", "x3", "v1", "v2", "v3"};
    for (int d = 0; d < 6; ++d) {
        const double emin = d < 3 ? 0.0 : -c->p.v_max, emax = d < 3 ? c->p.x_max[d] : c->p.v_max;
        const double de = (emax - emin) / (double)c->p.n[d];
        for (int i = 0; i < c->nw[d]; ++i) printf(" etas %s  %24.16E\n", names[d], emin + de * (double)(i + c->mn[d]));
    }
}
/* rho, Poisson, E for the current f, then one diagnostics row (interface.F90:186-231, 233-283) */
void sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics_init(void **sim) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics_init";
    Compat6d *c = self(sim, fun);
    push_mirror(c, fun);
    CK(sllb_sim6d_fields(c->S), fun);
    double row[14];
    CK(sllb_sim6d_diagnostics(c->S, (double)(c->first_time_step - 1) * c->p.delta_t, row), fun);
    write_row(c, row);
}
void sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics(void **sim, int32_t *time_step_number) {
    const char *fun = "sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics";
    Compat6d *c = self(sim, fun);
    push_mirror(c, fun);
    CK(sllb_sim6d_fields(c->S), fun);
    double row[14];
    CK(sllb_sim6d_diagnostics(c->S, (double)(*time_step_number) * c->p.delta_t, row), fun);
    write_row(c, row);
}

/* sll_s_check_diagnostics (sll_m_sim_6d_utilities.F90:648-687): 3 x 14 numbers, max abs difference < 5e-7.
 * Returns 0 when passed; prints like the reference. */
int sllb_sim6d_compat_check(const char *reffile, const char *simfile) {
    double a[42], b[42];
    FILE *fa = fopen(simfile, "r"), *fb = fopen(reffile, "r");
    if (!fa || !fb) {
        if (fa) fclose(fa);
        if (fb) fclose(fb);
        printf(" FAILED. (cannot open %s or %s)\n", simfile, reffile);
        return fail(SLLB_ERR_INVALID, "compat_check: cannot open the result or the reference file");
    }
    int na = 0, nb = 0;
    while (na < 42 && fscanf(fa, "%lf", &a[na]) == 1) ++na;
    while (nb < 42 && fscanf(fb, "%lf", &b[nb]) == 1) ++nb;
    fclose(fa); fclose(fb);
    if (na < 42 || nb < 42) { printf(" FAILED. (need 3 x 14 numbers)\n"); return fail(SLLB_ERR_INVALID, "compat_check: fewer than 3 x 14 numbers"); }
    double err = 0.0;
    for (int k = 0; k < 42; ++k) err = fmax(err, fabs(a[k] - b[k]));
    printf(" Max error in time history diagnostics:  %24.16E\n", err);
    if (err < 5.0e-7) { printf(" PASSED.\n"); return SLLB_OK; }
    printf(" FAILED.\n");
    return fail(SLLB_ERR_INVALID, "compat_check: diagnostics differ from the reference file");
}

} // extern "C"
