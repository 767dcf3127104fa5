// generate.cpp -- minimal compiled-language caller of the C ABI (include/lmrs_b200.h): the greedy generate loop of the
// reference's chat binary (src/bin/chat.rs:188-226 with --temperature 0) on token ids, without the tokenizer.
//   usage: generate <model.lmrs> [n_new_tokens] [first_token ...]
//   environment: LMRS_B200_GPUS=N    all N GPUs of this process behind one handle (lmrs_b200_create_multi)
//                GENERATE_SERIAL=1   the reference's own call pattern: one forward() per prompt token, logits to the host,
//                                    host argmax (chat.rs:196-214, sampler.rs:109-113)
// Maps the model file like the reference does (memmap2, chat.rs:60-65) and hands it to lmrs_b200_create.  Default call
// pattern = what the GPU path makes worthwhile in a caller: the prompt's embeddings go through ONE
// batched fill_kv_cache (the multimodal bins already do this, chat.rs:110-119; row-wise identical results for LLAMA / PHI),
// the last prompt token and every generated one through generate_greedy (pick and feedback on the device).
// There is no CPU fallback: without an sm_100 device the create call fails and this exits 1.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lmrs_b200.h"

static int die(const char* what) {
    std::fprintf(stderr, "generate: %s: %s\n", what, lmrs_b200_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <model.lmrs> [n_new_tokens] [first_token ...]\n", argv[0]);
        return 2;
    }
    const int n_new = argc > 2 ? std::atoi(argv[2]) : 16;
    std::vector<uint32_t> prompt;
    for (int i = 3; i < argc; i++) prompt.push_back((uint32_t)std::strtoul(argv[i], nullptr, 10));
    if (prompt.empty()) prompt.push_back(1);

    const int fd = open(argv[1], O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) { std::perror(argv[1]); return 2; }
    void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { std::perror("mmap"); return 2; }

    lmrs_b200_t* model = nullptr;
    size_t end_offset = 0;
    const char* gpus_env = std::getenv("LMRS_B200_GPUS");
    const int n_gpus = gpus_env ? std::atoi(gpus_env) : 1;
    if (n_gpus > 1) {
        if (lmrs_b200_create_multi(static_cast<const uint8_t*>(map), (size_t)st.st_size, n_gpus, &model, &end_offset)) return die("create_multi");
    } else if (lmrs_b200_create(static_cast<const uint8_t*>(map), (size_t)st.st_size, -1, &model, &end_offset)) return die("create");
    munmap(map, (size_t)st.st_size);   // every tensor now lives in HBM; the library keeps no pointer into the file
    close(fd);

    lmrs_args_t args;
    if (lmrs_b200_args(model, &args)) return die("args");
    std::fprintf(stderr, "%s v%s: dim %u, %u layers, vocab %u, q_type %u, model_type %u\n", argv[1], lmrs_b200_version(),
                 args.dim, args.n_layers, args.vocab_size, (unsigned)args.q_type, (unsigned)args.model_type);

    uint32_t pos = 0, token = prompt[0];
    float* logits = nullptr;
    auto argmax = [&]() {
        uint32_t best = 0;
        for (uint32_t i = 1; i < args.vocab_size; i++)
            if (logits[i] > logits[best]) best = i;
        return best;
    };
    for (uint32_t& t : prompt) t %= args.vocab_size;
    if (!std::getenv("GENERATE_SERIAL")) {
        // batched prompt: every token but the last through get_embeddings + fill_kv_cache, then one device-side greedy loop
        if (prompt.size() > 1) {
            std::vector<float> emb((prompt.size() - 1) * (size_t)args.dim);
            if (lmrs_b200_get_embeddings(model, prompt.data(), prompt.size() - 1, emb.data())) return die("get_embeddings");
            if (lmrs_b200_fill_kv_cache(model, emb.data(), emb.size(), 0, &pos)) return die("fill_kv_cache");
        }
        std::vector<uint32_t> out((size_t)(n_new > 0 ? n_new : 1));
        uint32_t n_out = 0;
        const auto tb = std::chrono::steady_clock::now();
        if (lmrs_b200_generate_greedy(model, prompt.back(), pos, (uint32_t)n_new, -1, out.data(), &n_out)) return die("generate_greedy");
        const double sb = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count();
        for (uint32_t i = 0; i < n_out; i++) std::printf("%u ", out[i]);
        std::printf("\n");
        std::fprintf(stderr, "%u tokens in %.3f s: %.1f tok/s (batched prompt, greedy pick and feedback on the device)\n", n_out, sb, n_out / sb);
        lmrs_b200_destroy(model);
        return 0;
    }
    for (size_t i = 0; i < prompt.size(); i++) {                     // prompt processing, one token per forward
        token = prompt[i];
        if (lmrs_b200_forward(model, token, pos++, &logits)) return die("forward");
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_new && pos < args.seq_len; i++) {
        token = argmax();
        std::printf("%u ", token);
        if (lmrs_b200_forward(model, token, pos++, &logits)) return die("forward");
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("\n");
    std::fprintf(stderr, "%d tokens in %.3f s: %.1f tok/s (end to end: logits to host, host argmax)\n", n_new, s, n_new / s);
    lmrs_b200_destroy(model);
    return 0;
}
