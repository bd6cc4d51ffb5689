// Host-side mirror of the reference's block interface, so that the GPU receive block can be dropped into
// fun_ofdm's receiver_chain (reference src/block.h:36-112, src/tagged_vector.h:25-94).  Same names,
// same members, same meaning: a maintainer who already includes the reference's own block.h /
// tagged_vector.h uses those instead (define B200_USE_REFERENCE_HEADERS before including b200_rx.h).
#ifndef B200_FUN_API_H
#define B200_FUN_API_H

#include <complex>
#include <string>
#include <vector>

#ifndef BUFFER_MAX
#define BUFFER_MAX 65536 // block.h:22
#endif

namespace fun
{
    // tagged_vector.h:25-34
    enum vector_tag { NONE, STS_START, STS_END, LTS_START, LTS1, LTS2, START_OF_FRAME };

    // tagged_vector.h:82-94: 24 bytes
    struct tagged_sample
    {
        std::complex<double> sample;
        vector_tag tag;
        tagged_sample() { tag = NONE; }
    };

    // tagged_vector.h:43-76
    template<int N>
    struct tagged_vector
    {
        std::complex<double> samples[N];
        vector_tag tag;
        tagged_vector(vector_tag _tag = NONE) { tag = _tag; }
    };

    // block.h:36-60
    class block_base
    {
    public:
        block_base(std::string block_name) : name(block_name) {}
        virtual ~block_base() {}
        virtual void work() = 0;
        std::string name;
    };

    // block.h:68-112
    template<typename I, typename O>
    class block : public block_base
    {
    public:
        block(std::string block_name) : block_base(block_name)
        {
            input_buffer.reserve(BUFFER_MAX);
            output_buffer.reserve(BUFFER_MAX);
        }
        virtual void work() = 0;
        std::vector<I> input_buffer;
        std::vector<O> output_buffer;
    };
}

#endif
