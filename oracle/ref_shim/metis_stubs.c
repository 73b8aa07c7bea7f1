/*
 * metis_stubs.c -- link-time placeholders (TEST INFRASTRUCTURE).
 * SuperLU_DIST 5.4.0's get_perm_c*.c reference METIS/ParMETIS orderings
 * (vendored config: HAVE_PARMETIS).  Saena sets options.ColPerm = NATURAL
 * (/root/reference/src/saena_object_solve.cpp:400-401), so they are never
 * called; abort loudly if that ever changes.
 */
#include <stdio.h>
#include <stdlib.h>
static void never(const char *n) { fprintf(stderr, "%s: METIS is not part of the oracle build\n", n); abort(); }
int METIS_NodeND(void) { never("METIS_NodeND"); return 0; }
int METIS_EdgeND(void) { never("METIS_EdgeND"); return 0; }
int ParMETIS_V3_NodeND(void) { never("ParMETIS_V3_NodeND"); return 0; }
