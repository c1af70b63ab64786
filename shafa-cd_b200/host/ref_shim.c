/*
 * ref_shim.c -- the one symbol the reference's driver still wants from utils/multithread.c when its modules
 * are replaced by host/mod_*.c: shafa.c:55-56 sets NO_MULTITHREAD for --no-multithread.  There are no per-block
 * host threads behind these modules (all blocks go through one batched GPU call), so the flag is accepted and
 * has nothing to switch.  Linked ONLY into a reference-driver build, never into ours.
 */
#include <stdbool.h>

bool NO_MULTITHREAD = false;
