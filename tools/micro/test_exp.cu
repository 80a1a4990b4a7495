// Host check of pyr::exp_tab (csrc/pyr_exp.cuh) against std::exp: max error in ulp.
//   nvcc -O2 -o /tmp/test_exp tools/micro/test_exp.cu && /tmp/test_exp
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>

#include "../../pyrate_b200/csrc/pyr_exp.cuh"

static double ulp_err(double got, long double want) {
    double w = (double)want;
    double ulp = std::nextafter(std::fabs(w), INFINITY) - std::fabs(w);
    return (double)(std::fabs((long double)got - want) / ulp);
}

int main() {
    std::mt19937_64 rng(12345);
    double worst = 0.0, worst_x = 0.0;
    const double ranges[4][2] = {{-700.0, 700.0}, {-30.0, 0.0}, {-1.0, 1.0}, {-1e-3, 1e-3}};
    long count = 0;
    for (int r = 0; r < 4; ++r) {
        std::uniform_real_distribution<double> u(ranges[r][0], ranges[r][1]);
        for (int i = 0; i < 2000000; ++i) {
            const double x = u(rng);
            const double e = ulp_err(pyr::exp_tab(x, pyr::kExp2Tab), expl((long double)x));
            if (e > worst) { worst = e; worst_x = x; }
            ++count;
        }
    }
    const double special[] = {0.0, -0.0, 1.0, -1.0, 0.021660849335603416 * 16, -699.9, 699.9, 1e-300, -1e-300};
    for (double x : special) {
        const double e = ulp_err(pyr::exp_tab(x, pyr::kExp2Tab), expl((long double)x));
        if (e > worst) { worst = e; worst_x = x; }
    }
    // arguments outside the table range: 0 below, NaN above and for NaN
    const bool range_ok = pyr::exp_tab_any(-700.5, pyr::kExp2Tab) == 0.0 && pyr::exp_tab_any(-1e300, pyr::kExp2Tab) == 0.0 &&
                          pyr::exp_tab_any(-INFINITY, pyr::kExp2Tab) == 0.0 && std::isnan(pyr::exp_tab_any(701.0, pyr::kExp2Tab)) &&
                          std::isnan(pyr::exp_tab_any(INFINITY, pyr::kExp2Tab)) && std::isnan(pyr::exp_tab_any(NAN, pyr::kExp2Tab)) &&
                          std::isnan(pyr::exp_tab_any(-NAN, pyr::kExp2Tab)) &&
                          pyr::exp_tab_any(-3.25, pyr::kExp2Tab) == pyr::exp_tab(-3.25, pyr::kExp2Tab);
    if (!range_ok) { printf("range handling of exp_tab_any failed\n"); return 2; }
    printf("samples %ld max_ulp_err %.4f at x = %.17g\n", count, worst, worst_x);
    return worst <= 1.01 ? 0 : 1;
}
