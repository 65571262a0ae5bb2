/* timing.c -- see timing.h */
#include "timing.h"

void get_timer_resolution(struct timespec* val) { clock_getres(CLOCK_REALTIME, val); }

void get_time(volatile struct timespec* val) { clock_gettime(CLOCK_REALTIME, (struct timespec*)val); }

double get_time_diff(struct timespec* start, struct timespec* finish)
{
	return (double)(finish->tv_sec - start->tv_sec) + 1e-9 * (double)(finish->tv_nsec - start->tv_nsec);
}
