/* timing.h -- wall-clock helpers of the b200 drivers: same interface as the reference's
 * <test>/timing.h (get_time / get_time_diff over CLOCK_REALTIME, timing.c:16-42), so the phase
 * columns t_init/t_alloc/t_load/t_comp/t_save/t_free mean the same thing.  Kernel time (t_krn)
 * is NOT taken with these: it is CUDA-event time reported by b200_run (b200_stats). */
#ifndef B200_TIMING_H
#define B200_TIMING_H

#include <time.h>

void get_timer_resolution(struct timespec* val);
void get_time(volatile struct timespec* val);
double get_time_diff(struct timespec* start, struct timespec* finish);

#endif
