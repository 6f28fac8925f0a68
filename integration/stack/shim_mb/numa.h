/* numa.h / numacompat1.h -- "no NUMA library here" stand-in for the reference's micro-benchmark driver (it only
 * binds when --numa is given and numa_available() >= 0). */
#ifndef STACK_SHIM_NUMA_H
#define STACK_SHIM_NUMA_H
typedef struct { unsigned long n[8]; } nodemask_t;
static inline int numa_available(void) { return -1; }
static inline int numa_max_node(void) { return 0; }
static inline void nodemask_zero(nodemask_t* m) { for (int i = 0; i < 8; i++) m->n[i] = 0; }
static inline void nodemask_set(nodemask_t* m, int node) { m->n[node / 64] |= 1ul << (node % 64); }
static inline void numa_bind(const nodemask_t*) {}
#endif
