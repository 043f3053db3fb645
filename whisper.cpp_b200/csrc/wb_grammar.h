// wb_grammar.h -- GBNF-constrained decoding (whisper_full_params.grammar_rules): a pushdown automaton over the rule table that
// penalises tokens the grammar cannot continue with and advances with every accepted token.
// Same observable behaviour as the whisper_grammar_* functions of src/whisper.cpp:5480-5923 (which are llama.cpp's grammar
// sampler): which token ids are penalised, and which parse states remain after a token.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "../../include/whisper_b200.h"

namespace wb {

struct Vocab;

struct Grammar {
    using Elem = whisper_grammar_element;
    using Stack = std::vector<const Elem *>;         // pending positions inside `rules`; back() is the next terminal to match
    // one END-terminated element list per rule; immutable and SHARED by all copies of a grammar (beam search copies decoders,
    // and the stacks hold pointers into this table)
    std::shared_ptr<const std::vector<std::vector<Elem>>> rules;
    std::vector<Stack> stacks;                       // every parse that is still alive
    uint32_t partial_value = 0; int partial_remain = 0;   // UTF-8 sequence left open by the last accepted token
    bool active() const { return rules && !rules->empty() && !stacks.empty(); }
};

Grammar grammar_init(const whisper_grammar_element * const * rules, size_t n_rules, size_t i_start_rule);     // whisper.cpp:5817-5855
// logits[id] -= penalty for every text token (id < eot, non-empty string) that no live parse can consume (whisper.cpp:5857-5899)
void grammar_penalize(const Vocab & vocab, const Grammar & g, float penalty, std::vector<float> & logits);
// advance all parses over the characters of `token` (special "[_...]" tokens leave the state alone; whisper.cpp:5901-5923)
void grammar_accept_token(const Vocab & vocab, Grammar & g, whisper_token token);

} // namespace wb
