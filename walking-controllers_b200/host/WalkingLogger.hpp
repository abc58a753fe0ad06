// Drop-in for the reference's WalkingLogger client (modules/Walking_module/include/WalkingLogger.hpp, WalkingLogger.tpp)
// together with the file writer of the WalkingLogger module it talks to
// (modules/WalkingLogger_module/src/WalkingLoggerModule.cpp:54-90 "record" -> header, :142-171 one row per vector):
// startRecord(names) opens the dataset and writes "time name ... "; sendData(vectors...) concatenates its arguments the
// way YarpHelper::sendVariadicVector does and writes "t v0 v1 ... " with the module's ostream default formatting.
// The YARP port / RPC hop between the two processes is not reproduced: rows go straight to the file.
// For batched runs the row itself is packed on the device by wc_glue_log_pack_dev (include/wc_capi.h).
#ifndef WC_HOST_WALKING_LOGGER_HPP
#define WC_HOST_WALKING_LOGGER_HPP

#include <fstream>
#include <initializer_list>
#include <string>
#include <vector>

#include "wc_compat.hpp"

class WalkingLogger {
    std::ofstream m_stream;
    size_t m_numberOfValues{0};
    double m_time{0.0}, m_dT{0.01};
    std::string m_fileName;

    void append(std::vector<double>&) {}
    template <typename T, typename... Args>
    void append(std::vector<double>& row, const T& v, const Args&... args)
    {
        for (size_t i = 0; i < static_cast<size_t>(v.size()); i++) row.push_back(v(static_cast<int>(i)));
        append(row, args...);
    }

public:
    // config: `sampling_time` advances the time column (the module uses wall-clock time); `file_name` overrides the
    // module's Dataset_<date>.txt
    bool configure(const yarp::os::Searchable& config, const std::string& name);
    bool startRecord(const std::initializer_list<std::string>& strings);
    void quit();
    bool writeRow(const double* values, size_t n); // one already-packed row (e.g. a column of wc_glue_log_pack_dev)
    const std::string& fileName() const { return m_fileName; }

    template <typename... Args>
    void sendData(const Args&... args)
    {
        std::vector<double> row;
        append(row, args...);
        writeRow(row.data(), row.size());
    }
};

#endif
