// See WalkingLogger.hpp.
#include "WalkingLogger.hpp"

#include <ctime>
#include <iomanip>
#include <iostream>
#include <sstream>

bool WalkingLogger::configure(const yarp::os::Searchable& config, const std::string& name)
{
    (void)name;
    if (config.isNull()) { std::cerr << "[configureLogger] Empty configuration for the force torque sensors.\n"; return false; }
    m_dT = wc_compat::cfg_number(config, "sampling_time", 0.01);
    m_fileName = config.check("file_name") ? wc_compat::cfg_string(config, "file_name", "") : std::string();
    return true;
}

bool WalkingLogger::startRecord(const std::initializer_list<std::string>& strings)
{
    if (m_stream.is_open()) { std::cerr << "[RPC Server] The stream is already open.\n"; return false; }   // LoggerModule.cpp:55-60
    auto it = strings.begin();
    if (it != strings.end() && *it == "record") ++it; // the first element is the RPC command (WalkingModule.cpp:1231)
    m_numberOfValues = 0;
    std::string head{"time "};                                                                               // :64-66
    for (; it != strings.end(); ++it) { head += *it + " "; m_numberOfValues++; }
    if (m_fileName.empty()) {                                                                                // :74-80
        std::time_t t = std::time(nullptr);
        std::tm tm = *std::localtime(&t);
        std::stringstream fileName;
        fileName << "Dataset_" << std::put_time(&tm, "%Y_%m_%d_%H_%M_%S") << ".txt";
        m_fileName = fileName.str();
    }
    m_stream.open(m_fileName.c_str());
    if (!m_stream.is_open()) return false;
    m_stream << head << std::endl;                                                                           // :85
    m_time = 0.0;
    return true;
}

void WalkingLogger::quit()
{
    if (m_stream.is_open()) m_stream.close();                                                                // :41-51
}

bool WalkingLogger::writeRow(const double* values, size_t n)
{
    if (!m_stream.is_open()) { std::cerr << "[updateModule] No stream is open. I cannot store your data.\n"; return false; } // :151-155
    if (n != m_numberOfValues) {                                                                             // :157-162
        std::cerr << "[updateModule] The size of the vector is not the one expected. Expected: " << m_numberOfValues
                  << " received: " << n << "\n";
        return false;
    }
    m_stream << m_time << " ";                                                                               // :165-170
    for (size_t i = 0; i < n; i++) m_stream << values[i] << " ";
    m_stream << std::endl;
    m_time += m_dT;
    return true;
}
