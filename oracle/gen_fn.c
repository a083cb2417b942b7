/* TEST INFRASTRUCTURE (oracle). Writes the N x N text matrix the reference programs read
 * on stdin (SC_128.c:149-158): Fn = F^{(x)n}, F=[[1,0],[1,1]], i.e. Fn[i][j]=1 iff (i&j)==j. */
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv)
{
    int nn = argc > 1 ? atoi(argv[1]) : 128;
    for (int i = 0; i < nn; i++) {
        for (int j = 0; j < nn; j++) { putchar(((i & j) == j) ? '1' : '0'); putchar(j + 1 < nn ? ' ' : '\n'); }
    }
    return 0;
}
