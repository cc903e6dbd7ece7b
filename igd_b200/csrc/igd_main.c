/* `igd` command: `igd search` runs on the GPU (igd_cli.c over the CUDA library); `igd create` is host-side C
 * (igd_create.c; north_star: "Index build stays host-side C").  Dispatch and help as src/igd.c:21-49. */
#include "igd_b200.h"

#include <stdio.h>
#include <string.h>

static int usage(int code)
{
    fprintf(stderr,
            "igd, v0.1.1\n"
            "usage:   igd <command> [options]\n"
            "         create    Create an igd database\n"
            "         search    Search an igd database\n");
    return code;
}

int main(int argc, char **argv)
{
    if (argc < 2) return usage(0);
    if (strcmp(argv[1], "create") == 0) return igd_create(argc, argv);
    if (strcmp(argv[1], "search") == 0) return igd_search(argc, argv);
    fprintf(stderr, "Unknown command\n");
    return usage(64);
}
