/* stand-in: the stencil operators the reference uses live in blitz/array.h of saras_b200/host/compat */
#include <blitz/array.h>
