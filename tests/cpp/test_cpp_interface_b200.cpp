// Drives the 6D simulation through the C interface of the reference's simulation
// (simulations/parallel/bsl_vp_3d3v_cart_dd/test_cpp_interface.cpp does the same against the Fortran build):
// init from a namelist file, query the local size, take the distribution, write it back, run, delete.
// Then the <prefix>.dat it wrote is checked against the golden file like the reference's ctest does.
// usage: test_cpp_interface_b200 <namelist> <reference .dat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>

#include "sll_b200_sim6d_compat.h"

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <namelist> <reffile>\n", argv[0]); return 2; }
    sll_s_allocate_collective();
    int fake_comm = 0;
    sll_s_set_communicator_collective(&fake_comm);

    void *sim = nullptr;
    std::printf("init\n");
    sim_bsl_vp_3d3v_cart_dd_slim_init(&sim, argv[1]);

    int32_t local_size[6];
    sim_bsl_vp_3d3v_cart_dd_slim_get_local_size(&sim, local_size);
    std::printf("local size [%d, %d, %d, %d, %d, %d]\n", local_size[0], local_size[1], local_size[2], local_size[3],
                local_size[4], local_size[5]);
    const size_t n = std::accumulate(local_size, local_size + 6, (size_t)1, std::multiplies<size_t>());

    // take the distribution, copy it out and back in (what a coupling framework does between runs)
    double *field = nullptr;
    sim_bsl_vp_3d3v_cart_dd_slim_get_distribution(&sim, &field);
    std::vector<double> copy(field, field + n);
    double mass = 0.0;
    for (double v : copy) mass += v;
    std::printf("sum f = %.15e\n", mass);
    std::memcpy(field, copy.data(), n * sizeof(double));

    std::printf("run\n");
    sim_bsl_vp_3d3v_cart_dd_slim_run(&sim);

    // a second hand-out after the run must reflect the advanced state
    sim_bsl_vp_3d3v_cart_dd_slim_get_distribution(&sim, &field);
    double diff = 0.0;
    for (size_t i = 0; i < n; ++i) { double d = field[i] - copy[i]; diff += d * d; }
    if (!(diff > 0.0)) { std::printf("FAILED: distribution unchanged by run\n"); return 1; }
    sim_bsl_vp_3d3v_cart_dd_slim_set_distribution(&sim, field);

    std::printf("delete\n");
    sim_bsl_vp_3d3v_cart_dd_slim_delete(&sim);
    sll_s_halt_collective();

    if (sllb_sim6d_compat_check(argv[2], "vp_3d3v_dd_b200.dat") != 0) return 1;
    std::printf("works in cpp\n");
    return 0;
}
