// hvb_prelude.cpp -- the device-side helpers (plan_format.h + hvb_device.cuh) as text, for the
// kernels NVRTC compiles at run time; build/hvb_prelude.inc is produced by the Makefile.
#include "hvb_gen.h"

const char* hvb_gen_prelude() {
    static const char* text =
#include "build/hvb_prelude.inc"
        ;
    return text;
}
