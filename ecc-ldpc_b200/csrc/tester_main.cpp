// ldpc_b200_tester -- command line equivalent of the reference's `ecc-ldpc` executable for the decoders of this
// library (main/Main.hs:38-48: "ecc-ldpc <EbN0...> <code-name...> [-m<N>]", e.g. NOTES.txt:2 `-m256`).
// For every code name and every Eb/N0 it runs N frames through ldpc_b200_ecc_run and prints the tester's line
//   <secs> <name> <EbN0> <#frames> <#bit-errors> <BER> +<p>% -<m>% 95%          (NOTES.txt:3)
// followed by the frame error rate and average iteration count as a comment.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/ldpc_b200_ecc.h"

static bool is_number(const char *s)
{
    char *end = nullptr;
    strtod(s, &end);
    return end && *end == 0 && end != s;
}

int main(int argc, char **argv)
{
    std::vector<double> ebn0s;
    std::vector<std::string> names;
    unsigned long long frames = 256, seed = 1;
    std::string codes = "codes";
    int random_messages = 0, max_batch = 16384;
    for (int i = 1; i < argc; i++) {
        std::string a(argv[i]);
        if (a.rfind("-m", 0) == 0 && a.size() > 2) frames = strtoull(a.c_str() + 2, nullptr, 10);
        else if (a == "--codes" && i + 1 < argc) codes = argv[++i];
        else if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (a == "--batch" && i + 1 < argc) max_batch = atoi(argv[++i]);
        else if (a == "--random-messages") random_messages = 1;
        else if (is_number(argv[i])) ebn0s.push_back(atof(argv[i]));
        else names.push_back(a);
    }
    if (ebn0s.empty() || names.empty()) {
        fprintf(stderr, "usage: %s [--codes DIR] [--seed S] [--batch B] [--random-messages] <EbN0...> <code-name...> [-m<N>]\n"
                        "  code-name: ldpc/<b200-minsum|b200-spa|b200-minsum-fp64|b200-spa-fp64>/<matrix-name>/<max-rounds>[/codeword/message]\n", argv[0]);
        return 2;
    }
    int rc = ldpc_b200_init(-1);
    if (rc) { fprintf(stderr, "%s\n", ldpc_b200_last_error()); return 1; }
    for (const std::string &nm : names) {
        ldpc_b200_ecc_t *e = nullptr;
        rc = ldpc_b200_ecc_open(codes.c_str(), nm.c_str(), max_batch, &e);
        if (rc == LDPC_B200_ENOMATCH) { fprintf(stderr, "no code matches %s\n", nm.c_str()); continue; }
        if (rc) { fprintf(stderr, "%s: %s\n", nm.c_str(), ldpc_b200_last_error()); return 1; }
        for (double eb : ebn0s) {
            ldpc_b200_counters_t c;
            memset(&c, 0, sizeof c);
            auto t0 = std::chrono::steady_clock::now();
            rc = ldpc_b200_ecc_run(e, eb, seed, 0, frames, random_messages, &c);
            if (rc) { fprintf(stderr, "%s: %s\n", nm.c_str(), ldpc_b200_last_error()); return 1; }
            double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            char line[512];
            ldpc_b200_ber_line(line, sizeof line, secs, ldpc_b200_ecc_name(e), eb, c.frames, c.bit_errors, ldpc_b200_ecc_message_length(e));
            printf("%s   # FER %.3e (%llu/%llu) avg-iters %.2f unconverged %llu kernel-ms %.2f\n", line,
                   c.frames ? (double)c.frame_errors / (double)c.frames : 0.0, (unsigned long long)c.frame_errors,
                   (unsigned long long)c.frames, c.frames ? (double)c.sum_iters / (double)c.frames : 0.0,
                   (unsigned long long)c.unconverged, c.kernel_ms);
            fflush(stdout);
        }
        ldpc_b200_ecc_close(e);
    }
    ldpc_b200_shutdown();
    return 0;
}
