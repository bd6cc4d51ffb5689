/* TEST INFRASTRUCTURE — build shim, not product code.
 *
 * Stand-in for boost::crc_32_type (Boost >= 1.59 asked for by the reference's
 * CMakeLists.txt:71, unpinned, not vendored).  boost::crc_32_type is
 * crc_optimal<32, 0x04C11DB7, 0xFFFFFFFF, 0xFFFFFFFF, true, true>, i.e. the
 * published CRC-32/ISO-HDLC (zlib) algorithm; check value
 * crc("123456789") = 0xCBF43926.  Byte-table driven like Boost's own.
 * Only the members the reference calls are provided
 * (ppdu.cpp:134-136, 267-269): process_bytes(), checksum().
 */
#ifndef B200RX_SHIM_BOOST_CRC_HPP
#define B200RX_SHIM_BOOST_CRC_HPP

#include <cstddef>
#include <cstdint>

namespace boost
{
    class crc_32_type
    {
    public:
        crc_32_type() : m_rem(0xFFFFFFFFu) {}

        void process_bytes(const void *buffer, std::size_t byte_count)
        {
            const std::uint32_t *tab = table();
            const unsigned char *b = static_cast<const unsigned char *>(buffer);
            std::uint32_t r = m_rem;
            for (std::size_t i = 0; i < byte_count; i++) r = tab[(r ^ b[i]) & 0xFFu] ^ (r >> 8);
            m_rem = r;
        }

        std::uint32_t checksum() const { return m_rem ^ 0xFFFFFFFFu; }

        void reset() { m_rem = 0xFFFFFFFFu; }

    private:
        std::uint32_t m_rem;

        static const std::uint32_t *table()
        {
            struct tab_t {
                std::uint32_t t[256];
                tab_t()
                {
                    for (std::uint32_t i = 0; i < 256; i++) {
                        std::uint32_t c = i;
                        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
                        t[i] = c;
                    }
                }
            };
            static const tab_t tab; /* thread-safe static init (C++11) */
            return tab.t;
        }
    };
}

#endif
