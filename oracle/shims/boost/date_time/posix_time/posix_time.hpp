/* TEST INFRASTRUCTURE — build shim, not product code.
 *
 * The reference uses Boost.DateTime only to time block->work()
 * (receiver_chain.cpp:84-88): ptime, microsec_clock::local_time(),
 * time_duration::total_microseconds().  std::chrono equivalent.
 */
#ifndef B200RX_SHIM_BOOST_POSIX_TIME_HPP
#define B200RX_SHIM_BOOST_POSIX_TIME_HPP

#include <chrono>

namespace boost { namespace posix_time {

    class time_duration
    {
    public:
        explicit time_duration(long long us = 0) : m_us(us) {}
        long long total_microseconds() const { return m_us; }
        long long total_milliseconds() const { return m_us / 1000; }
    private:
        long long m_us;
    };

    class ptime
    {
    public:
        ptime() : m_tp() {}
        explicit ptime(std::chrono::steady_clock::time_point tp) : m_tp(tp) {}
        time_duration operator-(const ptime &o) const
        {
            return time_duration(std::chrono::duration_cast<std::chrono::microseconds>(m_tp - o.m_tp).count());
        }
    private:
        std::chrono::steady_clock::time_point m_tp;
    };

    struct microsec_clock
    {
        static ptime local_time() { return ptime(std::chrono::steady_clock::now()); }
    };

}}

#endif
