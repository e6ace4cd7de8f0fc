/* integration/case_fork_main.c -- main() for the reference's UNMODIFIED test programs in the drop-in proof (integration/Makefile).
 *
 * The reference's own main (test/case_main.h) runs every TEST_CASE in one process and dies with the first failed `assert`, and it
 * selects cases by substring only.  For a per-case table on a GPU box this main keeps the reference's registration mechanism (the
 * `case_data` ELF section that TEST_CASE fills, test/case.h:66-95) but runs the cases in a forked child: when the child dies in
 * case k (assert / signal / per-case alarm) the parent records `CRASH k` and forks a new child that resumes at k + 1, so one
 * crash costs one extra CUDA context and never hides the cases behind it.  The parent itself never touches CUDA.
 *
 * Output protocol (parsed by integration/run_reference_tests.py): the case's own output (REQUIRE_... messages) followed by
 *   @@ PASS|FAIL|SKIP|CRASH <name>
 * and a final `@@ TOTAL pass skip fail crash`.  argv[1] (optional): substring filter, as in the reference's main;
 * argv[2] (optional): per-case time limit in seconds (default 180). */
#define CASE_DISABLE_MAIN
#include "case.h"
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>

extern case_t __start_case_data[];
extern case_t __stop_case_data[];

#define MAX_CASES 4096
static case_t* cases[MAX_CASES];

int main(int argc, char** argv)
{
	static const uint64_t the_sig = 0x883253372849284BULL; /* test/case.h: sig_head, sig_tail = sig_head + 2 */
	const char* const match = argc >= 2 && argv[1][0] ? argv[1] : 0;
	const int limit = argc >= 3 ? atoi(argv[2]) : 180;
	int n = 0;
	unsigned char* p;
	for (p = (unsigned char*)__start_case_data; p + sizeof(case_t) <= (unsigned char*)__stop_case_data && n < MAX_CASES; p += 8)
	{
		case_t* const c = (case_t*)p;
		if (c->sig_head == the_sig && c->sig_tail == the_sig + 2 && (!match || strstr(c->name, match)))
			cases[n++] = c;
	}
	setvbuf(stdout, 0, _IONBF, 0);
	int pass = 0, skip = 0, fail = 0, crash = 0, next = 0;
	while (next < n)
	{
		int fd[2];
		if (pipe(fd) != 0)
			return 2;
		const pid_t pid = fork();
		if (pid == 0)
		{
			close(fd[0]);
			if (__test_case_setup)
				__test_case_setup();
			int i;
			for (i = next; i < n; i++)
			{
				char msg = 'R';
				if (write(fd[1], &msg, 1) != 1)
					_exit(3);
				alarm(limit);
				int result = 0;
				cases[i]->func(cases[i]->name, &result);
				alarm(0);
				msg = result == 0 ? 'P' : result == -2 ? 'S' : 'F';
				if (write(fd[1], &msg, 1) != 1)
					_exit(3);
			}
			if (__test_case_teardown)
				__test_case_teardown();
			_exit(0);
		}
		close(fd[1]);
		int running = 0;
		char msg;
		while (read(fd[0], &msg, 1) == 1)
		{
			if (msg == 'R')
			{
				running = 1;
				continue;
			}
			running = 0;
			printf("\n@@ %s %s\n", msg == 'P' ? "PASS" : msg == 'S' ? "SKIP" : "FAIL", cases[next]->name);
			if (msg == 'P') pass++; else if (msg == 'S') skip++; else fail++;
			next++;
		}
		close(fd[0]);
		int status = 0;
		waitpid(pid, &status, 0);
		if (running && next < n)
		{
			printf("\n@@ CRASH %s (%s %d)\n", cases[next]->name, WIFSIGNALED(status) ? "signal" : "exit", WIFSIGNALED(status) ? WTERMSIG(status) : WEXITSTATUS(status));
			crash++, next++;
		} else if (next < n && !(WIFEXITED(status) && WEXITSTATUS(status) == 0)) {
			/* died outside a case (setup / teardown): do not loop forever */
			printf("\n@@ CRASH %s (outside a case, status %d)\n", cases[next]->name, status);
			crash++, next++;
		}
	}
	printf("\n@@ TOTAL %d %d %d %d\n", pass, skip, fail, crash);
	return fail + crash > 0;
}
