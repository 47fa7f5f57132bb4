// Compile-and-link check of include/MultiRayCaster.hpp against libmv_b200.so (no device needed):
// creating a caster without a GPU must fail cleanly with a message, never fall back.
#include "MultiRayCaster.hpp"
#include <cstdio>
int main()
{
    mvb200::MultiRayCaster rc;
    const bool ok = rc.Init(64, 36, 32, 16, 1, 1);
    std::printf("init=%d err=%s\n", ok ? 1 : 0, mvb200::MultiRayCaster::LastError().c_str());
    return 0;
}
