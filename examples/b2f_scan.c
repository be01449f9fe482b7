/* b2f_scan.c -- the whole filterbank stage of base2fil for one scan, from plain C.
 *
 * What /root/reference/base2fil.sh:404-448 does with one process_vdif/digifil child per IF, FIFOs and splice,
 * through the C ABI of include/b2f.h:
 *
 *   b2f_scan [--device N] [--nchan 128] [--tscrunch 16] [--bw 32] [--freq-lsb0 1254] [--pol 2] [--nbit 8]
 *            [--start 0] [--nsec 0] [--source NAME] OUT.fil IF1.vdif IF2.vdif ...
 *
 * IF i (1-based) is LSB when odd, USB when even, centred at freq-lsb0 + (i-1)*bw (base2fil.sh:54,65,254,407-414).
 * Build:  gcc -std=c99 -Iinclude examples/b2f_scan.c -o b2f_scan -Lfrb-baseband_b200 -lb2f -Wl,-rpath,$PWD/frb-baseband_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2f.h"

static int pol_mode_of(int pol) { /* process_vdif.py:163-176 */
    switch (pol) {
        case 0: return B2F_POL_P0;
        case 1: return B2F_POL_P1;
        case 2: return B2F_POL_I;
        case 3: return B2F_POL_I2;
        case 4: return B2F_POL_COHERENCE;
        default: return -1;
    }
}

int main(int argc, char** argv) {
    b2f_params p;
    b2f_scan_io io;
    double bw = 32.0, f0 = 1254.0;
    int pol = 2, a = 1;
    const char* source = "unknown";
    memset(&p, 0, sizeof p);
    memset(&io, 0, sizeof io);
    p.struct_size = sizeof p;
    io.struct_size = sizeof io;
    p.nchan = 128; p.tscrunch = 16; p.out_nbit = 8; p.in_nbit = 2; p.frame_bytes = 8032; p.header_bytes = 32;
    p.mask_faults = 1; p.rescale_interval_s = 10.0;
    if (argc == 2 && !strcmp(argv[1], "--version")) {
        printf("libb2f %d.%d.%d, %d device(s)\n", b2f_version() / 10000, b2f_version() / 100 % 100, b2f_version() % 100,
               b2f_device_count());
        return 0;
    }
    for (; a + 1 < argc && !strncmp(argv[a], "--", 2); a += 2) {
        const char* k = argv[a] + 2;
        const char* v = argv[a + 1];
        if (!strcmp(k, "device")) p.device = atoi(v);
        else if (!strcmp(k, "nchan")) p.nchan = atoi(v);
        else if (!strcmp(k, "tscrunch")) p.tscrunch = atoi(v);
        else if (!strcmp(k, "bw")) bw = atof(v);
        else if (!strcmp(k, "freq-lsb0")) f0 = atof(v);
        else if (!strcmp(k, "pol")) pol = atoi(v);
        else if (!strcmp(k, "nbit")) p.out_nbit = atoi(v);
        else if (!strcmp(k, "start")) io.start_s = atof(v);
        else if (!strcmp(k, "nsec")) io.nsec = atof(v);
        else if (!strcmp(k, "source")) source = v;
        else { fprintf(stderr, "unknown option --%s\n", k); return 2; }
    }
    if (argc - a < 2 || argc - a - 1 > B2F_MAX_IF || pol_mode_of(pol) < 0) {
        fprintf(stderr, "usage: b2f_scan [options] OUT.fil IF1.vdif [IF2.vdif ...]   (see the header of %s)\n", __FILE__);
        return 2;
    }
    {
        const char* out = argv[a];
        const char* const* in = (const char* const*)&argv[a + 1];
        struct b2f_plan* plan = NULL;
        b2f_scan_result r;
        int i, order[B2F_MAX_IF], rc;
        p.nif = argc - a - 1;
        p.pol_mode = pol_mode_of(pol);
        for (i = 0; i < p.nif; ++i) {
            p.bw_mhz[i] = (i + 1) % 2 ? -bw : bw;
            p.freq_mhz[i] = f0 + i * bw;
            order[i] = p.nif - 1 - i;                      /* splice: highest sky frequency first (base2fil.sh:350,367) */
        }
        for (i = 0; i < p.nif; ++i) p.if_order[i] = order[i];
        io.source_name = source;
        if (b2f_plan_create(&p, &plan)) { fprintf(stderr, "b2f_scan: %s\n", b2f_last_error()); return 1; }
        rc = b2f_run_scan(plan, p.nif, in, out, &io, &r);
        if (rc) fprintf(stderr, "b2f_scan: %s\n", b2f_last_error());
        else
            fprintf(stderr, "b2f_scan: %d IFs x %.3f s -> %s: %lld samples x %d channels in %.3f s (%.1f x real time); "
                            "frames ok/invalid/fill/bad = %llu/%llu/%llu/%llu\n",
                    p.nif, r.seconds_of_data, out, (long long)r.rows, p.nif * p.nchan, r.wall_s,
                    r.wall_s > 0 ? r.seconds_of_data / r.wall_s : 0.0, (unsigned long long)r.counters.frames_ok,
                    (unsigned long long)r.counters.frames_invalid, (unsigned long long)r.counters.frames_with_fill,
                    (unsigned long long)r.counters.frames_badhdr);
        b2f_plan_destroy(plan);
        return rc ? 1 : 0;
    }
}
