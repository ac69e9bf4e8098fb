// host_cpu_check.cpp -- the parts of the C++ host mirror (swiftest_host.hpp) that need no GPU: pair counts of
// pl%flatten, symba_pl%set_renc, and the "no CPU fallback" rule of the context.  Run by tests/test_abi.py on the CPU box;
// prints one line per check and exits nonzero on the first failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "swiftest_host.hpp"

using namespace swiftest;

static void require(bool ok, const char *what)
{
    printf("%s %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) exit(1);
}

int main()
{
    // swiftest_util_flatten_eucl_plpl / symba_util_flatten_eucl_plpl counts (swiftest_util.f90:1110, symba_util.f90:202)
    symba_pl pl;
    pl.setup(108);
    for (int i = 0; i < 108; ++i) pl.Gmass[(size_t)i] = (i < 57) ? 1e-5 : 1e-7;  // sorted by mass, 57 above GMTINY
    swiftest_parameters param;
    pl.flatten(param);
    require(pl.nplm == 108 && pl.nplpl == 108LL * 107 / 2 && pl.nplplm == pl.nplpl, "flatten without GMTINY: every pair");
    param.lmtiny_pl = true;
    param.GMTINY = 2.1554e-6;
    pl.flatten(param);
    long long brute = 0;
    for (int i = 0; i < 108; ++i)
        for (int j = i + 1; j < 108; ++j) brute += (i < 57);
    require(pl.nplm == 57 && pl.nplplm == brute, "flatten with GMTINY: nplplm = pairs with i <= nplm");

    // symba_util_set_renc (symba_util.f90:245-267): rhill * RHSCALE * RSHELL**irec, by repeated multiplication
    for (int i = 0; i < 108; ++i) pl.rhill[(size_t)i] = 0.01 + 1e-4 * i;
    pl.set_renc(0);
    require(pl.renc[5] == pl.rhill[5] * 6.5, "set_renc(0) = 6.5 rhill");
    pl.set_renc(3);
    require(pl.renc[7] == pl.rhill[7] * 6.5 * (((1.0 * 0.48075) * 0.48075) * 0.48075), "set_renc(3) multiplies RSHELL three times");

    // the context refuses to exist without an sm_100 GPU: a failed device call is fatal, never a fallback
    bool created = false, threw = false;
    try {
        cuda_context c(0);
        created = true;
    } catch (const fatal_error &e) {
        threw = true;
        printf("     fatal_error: %s\n", e.what());
    }
    require(created != threw, created ? "context created (GPU present)" : "no GPU: cuda_context throws fatal_error");
    bool threw2 = false;
    try {
        cuda_context c(4096);
    } catch (const fatal_error &) {
        threw2 = true;
    }
    require(threw2, "a device index that does not exist is fatal");
    printf("HOST-CPU-CHECK-OK\n");
    return 0;
}
