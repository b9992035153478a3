/*
 * dropin_consumer.c -- a C consumer of the reference's C ABI, written for this repository's tests.
 *
 * It binds the engine the way the reference file demo does (/root/reference/demo/c/koala_demo_file.c:265-333: dlopen + the same
 * eleven dlsym'd symbols with the same prototypes) and drives the same loop (:466-521): feed frames while start < total + delay,
 * zero-pad the tail, drop the first `delay` output samples, cut to the input length, time only pv_koala_process and print
 * "Real time factor" = compute / audio (:526-527).  The reference demo itself does not travel to the GPU box, so this file is what
 * proves there that a plain-C dlopen consumer of include/pv_koala.h works against libpv_koala_b200.so unmodified.
 * I/O is raw little-endian int16 mono (no WAV library): argv = library model access_key device in.raw out.raw   |   library -z
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

typedef struct pv_koala pv_koala_t;
typedef int pv_status_t;

static void *must(void *lib, const char *name) {
    void *p = dlsym(lib, name);
    if (!p) {
        fprintf(stderr, "Failed to load '%s': %s\n", name, dlerror());
        exit(2);
    }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s library model access_key device in.raw out.raw | %s library -z\n", argv[0], argv[0]);
        return 2;
    }
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) {
        fprintf(stderr, "Failed to open library at '%s': %s\n", argv[1], dlerror());
        return 2;
    }
    const char *(*status_to_string)(pv_status_t) = must(lib, "pv_status_to_string");
    int32_t (*sample_rate)(void) = must(lib, "pv_sample_rate");
    pv_status_t (*init)(const char *, const char *, const char *, pv_koala_t **) = must(lib, "pv_koala_init");
    void (*del)(pv_koala_t *) = must(lib, "pv_koala_delete");
    pv_status_t (*process)(pv_koala_t *, const int16_t *, int16_t *) = must(lib, "pv_koala_process");
    pv_status_t (*delay_sample)(const pv_koala_t *, int32_t *) = must(lib, "pv_koala_delay_sample");
    int32_t (*frame_length)(void) = must(lib, "pv_koala_frame_length");
    const char *(*version)(void) = must(lib, "pv_koala_version");
    pv_status_t (*get_error_stack)(char ***, int32_t *) = must(lib, "pv_get_error_stack");
    void (*free_error_stack)(char **) = must(lib, "pv_free_error_stack");
    pv_status_t (*list_devices)(char ***, int32_t *) = must(lib, "pv_koala_list_hardware_devices");
    void (*free_devices)(char **, int32_t) = must(lib, "pv_koala_free_hardware_devices");

    if (strcmp(argv[2], "-z") == 0) {
        char **devs = NULL;
        int32_t n = 0;
        pv_status_t st = list_devices(&devs, &n);
        if (st != 0) {
            fprintf(stderr, "Failed to list devices with '%s'\n", status_to_string(st));
            return 1;
        }
        for (int32_t i = 0; i < n; i++) fprintf(stdout, "%s\n", devs[i]);
        free_devices(devs, n);
        dlclose(lib);
        return 0;
    }
    if (argc < 7) return 2;
    pv_koala_t *koala = NULL;
    pv_status_t st = init(argv[3], argv[2], argv[4], &koala);
    if (st != 0) {
        fprintf(stderr, "Failed to init with '%s'", status_to_string(st));
        char **stack = NULL;
        int32_t depth = 0;
        if (get_error_stack(&stack, &depth) == 0 && depth > 0) {
            fprintf(stderr, ":\n");
            for (int32_t i = 0; i < depth; i++) fprintf(stderr, "  [%d] %s\n", i, stack[i]);
            free_error_stack(stack);
        } else {
            fprintf(stderr, ".\n");
        }
        return 1;
    }
    fprintf(stdout, "V%s\n", version());
    int32_t delay = 0;
    if (delay_sample(koala, &delay) != 0) return 1;
    const int32_t fl = frame_length();
    FILE *fi = fopen(argv[5], "rb");
    if (!fi) {
        fprintf(stderr, "Failed to open '%s'\n", argv[5]);
        return 1;
    }
    fseek(fi, 0, SEEK_END);
    const long total = ftell(fi) / 2;
    fseek(fi, 0, SEEK_SET);
    int16_t *in = calloc((size_t) total + fl, 2), *out = calloc((size_t) total + delay + 2 * fl, 2);
    int16_t *frame = calloc(fl, 2), *enhanced = calloc(fl, 2);
    if (fread(in, 2, total, fi) != (size_t) total) return 1;
    fclose(fi);
    double compute = 0.0, audio = 0.0;
    long start = 0, written = 0;
    while (start < total + delay) {
        memset(frame, 0, (size_t) fl * 2);
        const long have = total - start > fl ? fl : (total - start > 0 ? total - start : 0);
        memcpy(frame, in + start, (size_t) have * 2);
        struct timeval t0, t1;
        gettimeofday(&t0, NULL);
        st = process(koala, frame, enhanced);
        gettimeofday(&t1, NULL);
        if (st != 0) {
            fprintf(stderr, "Failed to process with '%s'\n", status_to_string(st));
            return 1;
        }
        compute += (double) (t1.tv_sec - t0.tv_sec) + 1e-6 * (double) (t1.tv_usec - t0.tv_usec);
        audio += (double) fl / (double) sample_rate();
        /* drop the first `delay` samples, keep at most `total` */
        long skip = start < delay ? delay - start : 0;
        if (skip > fl) skip = fl;
        long keep = fl - skip;
        if (written + keep > total) keep = total - written;
        if (keep > 0) {
            memcpy(out + written, enhanced + skip, (size_t) keep * 2);
            written += keep;
        }
        start += fl;
    }
    FILE *fo = fopen(argv[6], "wb");
    if (!fo || fwrite(out, 2, written, fo) != (size_t) written) return 1;
    fclose(fo);
    fprintf(stdout, "Real time factor : %.3f\n", compute / audio);
    del(koala);
    dlclose(lib);
    return 0;
}
