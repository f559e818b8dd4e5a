/* main() of the drop-in binary: the reference's own main (main.c:95, renamed at compile time
 * with -Dmain=tagdust_main when the reference is built as a library). */
int tagdust_main(int argc, char* argv[]);
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
int main(int argc, char* argv[])
{
	struct timespec ts;
	int rc;
	if (getenv("TDG_VERBOSE")) {
		clock_gettime(CLOCK_REALTIME, &ts);
		fprintf(stderr, "tagdust_b200: main() starts at epoch %.3f\n", ts.tv_sec + 1e-9 * ts.tv_nsec);
	}
	rc = tagdust_main(argc, argv);
	if (getenv("TDG_VERBOSE")) {   /* what is left after this line is process teardown (CUDA context, page cache) */
		clock_gettime(CLOCK_REALTIME, &ts);
		fprintf(stderr, "tagdust_b200: main() returns at epoch %.3f\n", ts.tv_sec + 1e-9 * ts.tv_nsec);
	}
	return rc;
}
