/* `igd` command: only the search path is GPU-native here; `igd create` stays with the reference's
 * host-side C (north_star: "Index build stays host-side C"). */
#include "igd_b200.h"

#include <stdio.h>
#include <string.h>

static int usage(int code)
{
    fprintf(stderr,
            "igd, v0.1.1\n"
            "usage:   igd <command> [options]\n"
            "         search    Search an igd database (GPU)\n");
    return code;
}

int main(int argc, char **argv)
{
    if (argc < 2) return usage(0);
    if (strcmp(argv[1], "search") == 0) return igd_search(argc, argv);
    fprintf(stderr, "Unknown command\n");
    return usage(64);
}
