/*
 * ref_prelude.h -- force-included (-include) in front of the UNMODIFIED reference
 * sources when they are compiled into oracle/_ref/ (see oracle/Makefile).
 *
 * The reference's timestepper leaves the tails of its coarse velocity arrays
 * uninitialised (multigrid.cpp:152-157) and then reads
 * them.  On a fresh process those pages are zero, so the reference "works", but
 * a test process that calls timestepper repeatedly can be handed recycled heap
 * chunks.  Mapping malloc to calloc pins the in-practice behaviour (zero tails)
 * without touching a single reference source line.  tests/test_oracle_vs_ref.py
 * checks that a build WITHOUT this prelude, called once in a fresh process
 * (oracle/_ref/ref_fresh), produces the same bits.
 */
#ifndef MGB200_REF_PRELUDE_H
#define MGB200_REF_PRELUDE_H
#include <stdlib.h>
#define malloc(sz) calloc(1, (sz))
#endif
