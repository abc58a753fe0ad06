// ini_dump.cpp — prints the hot-path parameters parsed from the reference's shipped .ini files by the
// Searchable stand-in (tests/test_host_cpp.py compares them with wc_*_default_config).
#include <cstdio>
#include "wc_compat.hpp"

int main(int argc, char** argv)
{
    yarp::os::Searchable c;
    for (int i = 1; i < argc; i++) if (!c.fromIniFile(argv[i])) { std::fprintf(stderr, "cannot read %s\n", argv[i]); return 2; }
    const char* keys[] = {"sampling_time", "com_height", "controllerHorizon", "stateWeightTriplets", "inputWeightTriplets", "foot_size",
                          "initial_zmp_position", "convex_hull_tolerance", "useCoMAsConstraint", "neckWeightTriplets",
                          "additional_rotation", "jointRegularization", "jointRegularizationWeights", "jointRegularizationGains",
                          "k_posCom", "k_posFoot", "k_attFoot", "k_neck"};
    for (const char* k : keys) {
        std::printf("%s", k);
        for (double v : c.list(k)) std::printf(" %.17g", v);
        std::printf("\n");
    }
    return 0;
}
