/* freesasa_b200/csrc/workers.c — fork/join helper for the host rows (reading, tree building): plain pthreads created per
 * call; the work items are milliseconds long, a thread costs ~20 us. */
#include "host_internal.h"

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

#define MAX_WORKERS 16

/* Threads the host layer may use: FREESASA_B200_THREADS, else the online processors, at most 16. */
int fsb_hardware_threads(void)
{
    const char *env = getenv("FREESASA_B200_THREADS");
    long n = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    return n > MAX_WORKERS ? MAX_WORKERS : (int)n;
}

struct part {
    void (*fn)(int, int, void *);
    int index, count;
    void *arg;
};

static void *run_part(void *p)
{
    struct part *q = p;
    q->fn(q->index, q->count, q->arg);
    return NULL;
}

/* fn(k, n_parts, arg) for k = 0 .. n_parts-1, part 0 on the calling thread; parts whose thread cannot be created run on
 * the calling thread as well.  Returns when all are done. */
void fsb_parallel_run(int n_parts, void (*fn)(int, int, void *), void *arg)
{
    struct part part[MAX_WORKERS];
    pthread_t thread[MAX_WORKERS];
    int k, started = 0;
    if (n_parts > MAX_WORKERS) n_parts = MAX_WORKERS;
    if (n_parts < 1) n_parts = 1;
    for (k = 0; k < n_parts; ++k) {
        part[k].fn = fn;
        part[k].index = k;
        part[k].count = n_parts;
        part[k].arg = arg;
    }
    for (k = 1; k < n_parts; ++k) {
        if (pthread_create(&thread[k], NULL, run_part, &part[k]) != 0) break;
        ++started;
    }
    run_part(&part[0]);
    for (k = started + 1; k < n_parts; ++k) run_part(&part[k]);
    for (k = 1; k <= started; ++k) pthread_join(thread[k], NULL);
}
