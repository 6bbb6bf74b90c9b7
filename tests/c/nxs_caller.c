/*
 * A plain C caller of the public API in include/nxs.h, in the manner of the
 * reference's own CLI (ref src/utils/benchmark.c:180-215: open the instance,
 * open the index, nxs_index_search per query, walk the response).  Compiled
 * by the tests and by bench.py against include/nxs.h and libnxsearch.so --
 * the boundary exercised from C, the way a user of the reference links it.
 *
 *   nxs_caller BASEDIR INDEX ALGO LIMIT QUERIES.txt results
 *       one line per query: "n id:score id:score ..." (score as %.9g)
 *   nxs_caller BASEDIR INDEX ALGO LIMIT QUERIES.txt latency [WARMUP]
 *       every query once through nxs_index_search, one at a time; prints a
 *       JSON object with p50/p90/p99/mean microseconds and queries/s
 *
 * fuzzymatch is off in both modes (the synthetic queries are exact terms).
 */
#define _GNU_SOURCE
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nxs.h"

static double
now_us(void)
{
	struct timespec ts;

	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}

static int
cmp_double(const void *a, const void *b)
{
	const double x = *(const double *)a, y = *(const double *)b;

	return (x > y) - (x < y);
}

int
main(int argc, char **argv)
{
	nxs_t *nxs;
	nxs_index_t *idx;
	nxs_params_t *params;
	char **queries = NULL, *line = NULL;
	size_t n = 0, cap = 0, len = 0;
	FILE *fp;
	bool latency;

	if (argc < 7) {
		fprintf(stderr, "usage: %s BASEDIR INDEX ALGO LIMIT QUERIES.txt results|latency [WARMUP]\n", argv[0]);
		return 2;
	}
	latency = strcmp(argv[6], "latency") == 0;
	if ((fp = fopen(argv[5], "r")) == NULL) {
		perror(argv[5]);
		return 1;
	}
	while (getline(&line, &len, fp) > 0) {
		line[strcspn(line, "\n")] = '\0';
		if (n == cap) {
			cap = cap ? cap * 2 : 1024;
			queries = realloc(queries, cap * sizeof(char *));
		}
		queries[n++] = strdup(line);
	}
	fclose(fp);
	free(line);

	if ((nxs = nxs_open(argv[1])) == NULL) {
		fprintf(stderr, "nxs_open failed\n");
		return 1;
	}
	if ((idx = nxs_index_open(nxs, argv[2])) == NULL) {
		const char *msg = NULL;

		nxs_get_error(nxs, &msg);
		fprintf(stderr, "nxs_index_open: %s\n", msg ? msg : "?");
		return 1;
	}
	params = nxs_params_create();
	nxs_params_set_str(params, "algo", argv[3]);
	nxs_params_set_uint(params, "limit", strtoull(argv[4], NULL, 10));
	nxs_params_set_bool(params, "fuzzymatch", false);

	if (!latency) {
		for (size_t i = 0; i < n; i++) {
			nxs_resp_t *resp = nxs_index_search(idx, params, queries[i], strlen(queries[i]));
			nxs_doc_id_t id;
			float score;

			if (resp == NULL) {
				const char *msg = NULL;

				printf("error %d %s\n", (int)nxs_get_error(nxs, &msg), msg ? msg : "");
				continue;
			}
			printf("%u", nxs_resp_resultcount(resp));
			nxs_resp_iter_reset(resp);
			while (nxs_resp_iter_result(resp, &id, &score))
				printf(" %llu:%.9g", (unsigned long long)id, score);
			printf("\n");
			nxs_resp_release(resp);
		}
	} else {
		const size_t warm = argc > 7 ? strtoull(argv[7], NULL, 10) : 16;
		double *us = malloc(sizeof(double) * (n ? n : 1)), sum = 0, t_all;
		unsigned long long results = 0;
		size_t m = 0;

		for (size_t i = 0; i < warm && i < n; i++) {
			nxs_resp_t *resp = nxs_index_search(idx, params, queries[i], strlen(queries[i]));

			if (resp)
				nxs_resp_release(resp);
		}
		t_all = now_us();
		for (size_t i = 0; i < n; i++) {
			const double t0 = now_us();
			nxs_resp_t *resp = nxs_index_search(idx, params, queries[i], strlen(queries[i]));
			nxs_doc_id_t id;
			float score;

			if (resp == NULL)
				continue;
			nxs_resp_iter_reset(resp);
			while (nxs_resp_iter_result(resp, &id, &score))
				results++;
			nxs_resp_release(resp);
			us[m] = now_us() - t0;
			sum += us[m++];
		}
		t_all = now_us() - t_all;
		if (m == 0) {
			fprintf(stderr, "no query succeeded\n");
			return 1;
		}
		qsort(us, m, sizeof(double), cmp_double);
		printf("{\"queries\": %zu, \"results\": %llu, \"p50_us\": %.1f, \"p90_us\": %.1f, \"p99_us\": %.1f, "
		    "\"mean_us\": %.1f, \"max_us\": %.1f, \"serial_queries_per_s\": %.1f}\n", m, results,
		    us[m / 2], us[(size_t)(m * 0.9)], us[(size_t)(m * 0.99)], sum / m, us[m - 1],
		    m / (t_all / 1e6));
		free(us);
	}
	nxs_params_release(params);
	nxs_index_close(idx);
	nxs_close(nxs);
	for (size_t i = 0; i < n; i++)
		free(queries[i]);
	free(queries);
	return 0;
}
