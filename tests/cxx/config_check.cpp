// Prints what Config makes of an INI file plus command-line overrides; built twice by tests/test_config_cpu.py:
// against the reference's own src/config.h + src/config.cpp and against fdm_b200/cxx/fdm_compat_config.h.
#include <cstdio>
#include <string>
#ifdef USE_REFERENCE_CONFIG
#include "config.h"
#else
#include "fdm_compat_config.h"
#endif

int main(int argc, char** argv)
{
    Config c;
    c.open(argv[1]);
    c.rewrite(argc - 1, argv + 1);
    const char* ints[][2] = {{"ns", "nx"}, {"ns", "nz"}, {"ns", "steps"}, {"plot", "interval"}, {"plot", "vtk"}, {"other", "check"},
                             {"pre", "early"}, {"ns", "spaced"}};
    for (auto& k : ints) printf("int %s:%s = %d\n", k[0], k[1], c.get(k[0], k[1], -7));
    const char* dbls[][2] = {{"ns", "Re"}, {"ns", "dt"}, {"ns", "x1"}, {"ns", "missing"}, {"ns", "tabbed"}};
    for (auto& k : dbls) printf("double %s:%s = %.17g\n", k[0], k[1], c.get(k[0], k[1], 0.125));
    const char* strs[][2] = {{"solver", "datatype"}, {"st", "input"}, {"ns", "label"}, {"nosuch", "key"}};
    for (auto& k : strs) printf("string %s:%s = '%s'\n", k[0], k[1], c.get(k[0], k[1], std::string("dflt")).c_str());
    return 0;
}
