// Host build of jues.jl_b200/csrc/synth_element.h: prints generator values for index quadruples read from
// stdin so that tests/test_synth_element.py can compare them bit for bit with the numpy generator.
#include "../../jues.jl_b200/csrc/synth_element.h"

#include <cinttypes>
#include <cstdio>
#include <cstring>

int main() {
    unsigned long long seed;
    double scale;
    if (scanf("%llu %lf", &seed, &scale) != 2) return 2;
    unsigned mu, nu, lam, sig;
    while (scanf("%u %u %u %u", &mu, &nu, &lam, &sig) == 4) {
        const double v = jues::synth_value(jues::synth_pair32(mu, nu), jues::synth_pair32(lam, sig), seed, scale);
        uint64_t bits;
        memcpy(&bits, &v, 8);
        printf("%016" PRIx64 "\n", bits);
    }
    return 0;
}
