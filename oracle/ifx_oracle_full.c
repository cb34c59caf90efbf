/* ifx_oracle_full.c — UNPINNED stages (placeholder; filled in below in this round). */
#include "ifx_oracle.h"
