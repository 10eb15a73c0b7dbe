// host-side probe of tc_layout() / the stage-ring allocation rule of tau_group_tc_kernel.cuh (tests/test_tc_layout_cpu.py)
#include <cstdio>
#include <cstdlib>
#include "../desman_b200/csrc/tau_group_tc_kernel.cuh"

// the shape rule of engine.cu (tc_shape): K blocks of <= 64, 32, 16 samples until the layout fits
static bool shape(int S, int G, int *SK, int *nkb, int *NC)
{
    *NC = (3 * G + 7) & ~7;
    for (int kmax = 64; kmax >= 16; kmax >>= 1) {
        const int nb = (S + kmax - 1) / kmax;
        *nkb = nb;
        *SK = (((S + nb - 1) / nb) + 3) & ~3;
        if (tc_layout(S, G, *SK, *nkb, *NC).total + 2048 <= 227 * 1024) return true;
    }
    return false;
}

int main()
{
    // 1. layouts: every shape the engine may pick
    for (int G = 1; G <= 21; G++)
        for (int S = 1; S <= 512; S += (S < 70 ? 1 : 13)) {
            int SK, nkb, NC;
            if (!shape(S, G, &SK, &nkb, &NC)) { printf("nofit %d %d\n", S, G); continue; }
            const TcLayout L = tc_layout(S, G, SK, nkb, NC);
            printf("layout %d %d %d %d %d %zu %d %d %d %zu %zu %d\n", S, G, SK, nkb, NC, L.total, L.nst, L.ntb, L.nacc, L.stage_bytes,
                   L.table_bytes, L.tmem_cols);
        }
    for (int k = 0; k < 3; k++) {                                       // the BASELINE shapes C3, C4, C5
        const int Ss[3] = {64, 256, 128}, Gs[3] = {8, 16, 20};
        int SK, nkb, NC;
        printf("baseline %d %d %d\n", Ss[k], Gs[k], shape(Ss[k], Gs[k], &SK, &nkb, &NC) ? 1 : 0);
    }
    // 2. the byte ring of the copy issuer, replayed on the host with random item sizes: no live allocation is overwritten
    srand(12345);
    for (int trial = 0; trial < 200; trial++) {
        const unsigned F = 4096u * (1 + rand() % 16), R = F * (2 + rand() % 3);
        unsigned start[TC_MAXRING], end[TC_MAXRING], u = 0, u_tail = 0, head = 0;
        for (int it = 0; it < 400; it++, u++) {
            const unsigned bytes = 4096u * (1 + rand() % (F / 4096u));
            if (head + F > R) head = 0;
            while (u_tail < u) {
                const unsigned s = u_tail % TC_MAXRING;
                const bool full = (u - u_tail) == TC_MAXRING, overlap = start[s] < head + bytes && head < end[s];
                if (!full && !overlap) break;
                u_tail++;                                               // (the device waits for the MMA commit of that allocation here)
            }
            for (unsigned w = u_tail; w < u; w++) {                     // every allocation still live must be disjoint from the new one
                const unsigned s = w % TC_MAXRING;
                if (start[s] < head + bytes && head < end[s]) { printf("ring OVERLAP trial %d item %d\n", trial, it); return 1; }
            }
            if (head + bytes > R || head + F > R) { printf("ring OUT OF RANGE trial %d item %d\n", trial, it); return 1; }
            start[u % TC_MAXRING] = head; end[u % TC_MAXRING] = head + bytes;
            head += bytes;
            if (rand() % 3 == 0 && u_tail < u) u_tail++;                // releases arrive in order, at random times
        }
    }
    printf("ring ok\n");
    return 0;
}
