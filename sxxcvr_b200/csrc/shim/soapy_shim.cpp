// Implementation of the minimal SoapySDR-compatible shim: registry, factory, logger,
// tick/time conversion and the benign Device defaults.  Host-side plumbing only; no sample
// ever passes through this file.
#include <SoapySDR/Device.hpp>
#include <SoapySDR/Formats.h>
#include <SoapySDR/Logger.hpp>
#include <SoapySDR/Registry.hpp>
#include <SoapySDR/Time.hpp>

#include "../sx_time.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <sstream>
#include <stdexcept>

// ---------------------------------------------------------------------------------------
// Registry
// ---------------------------------------------------------------------------------------
namespace {
struct DriverEntry {
    SoapySDR::FindFunction find;
    SoapySDR::MakeFunction make;
};
// Function-local statics: driver modules register from static initialisers, which may run
// before this translation unit's own globals are constructed.
std::map<std::string, DriverEntry> &driverTable(void)
{
    static std::map<std::string, DriverEntry> table;
    return table;
}
std::mutex &driverMutex(void)
{
    static std::mutex m;
    return m;
}
}

SoapySDR::Registry::Registry(const std::string &name, const FindFunction &find,
                             const MakeFunction &make, const std::string &abi)
{
    if (abi != SOAPY_SDR_ABI_VERSION) {
        std::fprintf(stderr, "SoapySDR shim: driver %s built for ABI %s, expected %s\n",
                     name.c_str(), abi.c_str(), SOAPY_SDR_ABI_VERSION);
        return;
    }
    std::lock_guard<std::mutex> lock(driverMutex());
    if (driverTable().count(name) != 0) {
        std::fprintf(stderr, "SoapySDR shim: driver %s registered twice, keeping the first\n",
                     name.c_str());
        return;
    }
    driverTable()[name] = DriverEntry{find, make};
    _name = name;
}

SoapySDR::Registry::~Registry(void)
{
    if (_name.empty())
        return;
    std::lock_guard<std::mutex> lock(driverMutex());
    driverTable().erase(_name);
}

std::vector<std::string> SoapySDR::Registry::listDrivers(void)
{
    std::lock_guard<std::mutex> lock(driverMutex());
    std::vector<std::string> names;
    for (const auto &kv : driverTable())
        names.push_back(kv.first);
    return names;
}

SoapySDR::FindFunctions SoapySDR::Registry::listFindFunctions(void)
{
    std::lock_guard<std::mutex> lock(driverMutex());
    FindFunctions out;
    for (const auto &kv : driverTable())
        out[kv.first] = kv.second.find;
    return out;
}

SoapySDR::MakeFunctions SoapySDR::Registry::listMakeFunctions(void)
{
    std::lock_guard<std::mutex> lock(driverMutex());
    MakeFunctions out;
    for (const auto &kv : driverTable())
        out[kv.first] = kv.second.make;
    return out;
}

// ---------------------------------------------------------------------------------------
// Kwargs markup: "key=value, key2=value2"
// ---------------------------------------------------------------------------------------
static std::string trimmed(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t");
    if (a == std::string::npos)
        return "";
    size_t b = s.find_last_not_of(" \t");
    return s.substr(a, b - a + 1);
}

SoapySDR::Kwargs SoapySDR::KwargsFromString(const std::string &markup)
{
    Kwargs out;
    std::stringstream ss(markup);
    std::string item;
    while (std::getline(ss, item, ',')) {
        size_t eq = item.find('=');
        std::string key = trimmed(eq == std::string::npos ? item : item.substr(0, eq));
        std::string val = eq == std::string::npos ? "" : trimmed(item.substr(eq + 1));
        if (!key.empty())
            out[key] = val;
    }
    return out;
}

std::string SoapySDR::KwargsToString(const Kwargs &args)
{
    std::string out;
    for (const auto &kv : args) {
        if (!out.empty())
            out += ", ";
        out += kv.first + "=" + kv.second;
    }
    return out;
}

// ---------------------------------------------------------------------------------------
// Factory
// ---------------------------------------------------------------------------------------
SoapySDR::KwargsList SoapySDR::Device::enumerate(const Kwargs &args)
{
    KwargsList found;
    const bool filtered = args.count("driver") != 0;
    for (const auto &kv : Registry::listFindFunctions()) {
        if (filtered && args.at("driver") != kv.first)
            continue;
        for (auto result : kv.second(args)) {
            result["driver"] = kv.first;
            found.push_back(result);
        }
    }
    return found;
}

SoapySDR::KwargsList SoapySDR::Device::enumerate(const std::string &args)
{
    return enumerate(KwargsFromString(args));
}

SoapySDR::Device *SoapySDR::Device::make(const Kwargs &inputArgs)
{
    KwargsList found = enumerate(inputArgs);
    if (found.empty())
        throw std::runtime_error("SoapySDR::Device::make() no match");
    Kwargs args = found.front();
    for (const auto &kv : inputArgs) // caller's keys win, discovered keys fill in
        args[kv.first] = kv.second;
    MakeFunctions makers = Registry::listMakeFunctions();
    auto it = makers.find(args["driver"]);
    if (it == makers.end())
        throw std::runtime_error("SoapySDR::Device::make() no driver " + args["driver"]);
    return it->second(args);
}

SoapySDR::Device *SoapySDR::Device::make(const std::string &args)
{
    return make(KwargsFromString(args));
}

void SoapySDR::Device::unmake(Device *device)
{
    delete device;
}

// ---------------------------------------------------------------------------------------
// Device defaults
// ---------------------------------------------------------------------------------------
namespace SoapySDR {
Device::~Device(void) {}
std::string Device::getDriverKey(void) const { return ""; }
std::string Device::getHardwareKey(void) const { return ""; }
Kwargs Device::getHardwareInfo(void) const { return Kwargs(); }
size_t Device::getNumChannels(const int) const { return 0; }
std::vector<std::string> Device::getStreamFormats(const int, const size_t) const
{
    return std::vector<std::string>();
}
std::string Device::getNativeStreamFormat(const int, const size_t, double &fullScale) const
{
    fullScale = double(1 << 15);
    return SOAPY_SDR_CS16;
}
Stream *Device::setupStream(const int, const std::string &, const std::vector<size_t> &,
                            const Kwargs &)
{
    return nullptr;
}
void Device::closeStream(Stream *) {}
size_t Device::getStreamMTU(Stream *) const { return 1024; }
int Device::activateStream(Stream *, const int flags, const long long, const size_t)
{
    return (flags == 0) ? 0 : SOAPY_SDR_NOT_SUPPORTED;
}
int Device::deactivateStream(Stream *, const int flags, const long long)
{
    return (flags == 0) ? 0 : SOAPY_SDR_NOT_SUPPORTED;
}
int Device::readStream(Stream *, void *const *, const size_t, int &, long long &, const long)
{
    return SOAPY_SDR_NOT_SUPPORTED;
}
int Device::writeStream(Stream *, const void *const *, const size_t, int &, const long long,
                        const long)
{
    return SOAPY_SDR_NOT_SUPPORTED;
}
std::vector<std::string> Device::listAntennas(const int, const size_t) const
{
    return std::vector<std::string>();
}
void Device::setAntenna(const int, const size_t, const std::string &) {}
std::string Device::getAntenna(const int, const size_t) const { return ""; }
std::vector<std::string> Device::listGains(const int, const size_t) const
{
    return std::vector<std::string>();
}
void Device::setGain(const int, const size_t, const double) {}
void Device::setGain(const int, const size_t, const std::string &, const double) {}
// Overall gain, as upstream's default does it: the elements' gains above their minima, summed,
// on top of the overall minimum.  (The SoapySX driver only implements the named elements,
// reference SoapySX.cpp:1279-1394, and relies on these defaults for the overall figures.)
double Device::getGain(const int direction, const size_t channel) const
{
    double gain = 0.0;
    for (const auto &name : listGains(direction, channel))
        gain += getGain(direction, channel, name) - getGainRange(direction, channel, name).minimum();
    return gain + getGainRange(direction, channel).minimum();
}
double Device::getGain(const int, const size_t, const std::string &) const { return 0.0; }
Range Device::getGainRange(const int direction, const size_t channel) const
{
    double lowest = 0.0, span = 0.0;
    for (const auto &name : listGains(direction, channel)) {
        const Range r = getGainRange(direction, channel, name);
        lowest += r.minimum();
        span += r.maximum() - r.minimum();
    }
    return Range(lowest, lowest + span);
}
Range Device::getGainRange(const int, const size_t, const std::string &) const
{
    return Range(0.0, 0.0);
}
void Device::setFrequency(const int, const size_t, const double, const Kwargs &) {}
double Device::getFrequency(const int, const size_t) const { return 0.0; }
void Device::setSampleRate(const int, const size_t, const double) {}
double Device::getSampleRate(const int, const size_t) const { return 0.0; }
std::vector<double> Device::listSampleRates(const int, const size_t) const
{
    return std::vector<double>();
}
RangeList Device::getSampleRateRange(const int, const size_t) const { return RangeList(); }
bool Device::hasHardwareTime(const std::string &) const { return false; }
long long Device::getHardwareTime(const std::string &) const { return 0; }
void Device::writeRegister(const std::string &, const unsigned, const unsigned) {}
unsigned Device::readRegister(const std::string &, const unsigned) const { return 0; }
void Device::writeRegisters(const std::string &name, const unsigned addr,
                            const std::vector<unsigned> &value)
{
    for (size_t i = 0; i < value.size(); i++)
        writeRegister(name, addr + unsigned(i), value[i]);
}
std::vector<unsigned> Device::readRegisters(const std::string &name, const unsigned addr,
                                            const size_t length) const
{
    std::vector<unsigned> out(length);
    for (size_t i = 0; i < length; i++)
        out[i] = readRegister(name, addr + unsigned(i));
    return out;
}
void Device::writeSetting(const std::string &, const std::string &) {}
std::string Device::readSetting(const std::string &) const { return ""; }
}

// ---------------------------------------------------------------------------------------
// Logger.  The level gate comes before any formatting: the stream path logs on every
// read/write at DEBUG (reference SoapySX.cpp:904, :996) and must not pay for it.
// ---------------------------------------------------------------------------------------
static SoapySDRLogLevel g_logLevel = SOAPY_SDR_INFO;
static SoapySDRLogHandler g_logHandler = nullptr;

static void defaultLogHandler(const SoapySDRLogLevel level, const char *message)
{
    static const char *names[] = {"", "FATAL", "CRITICAL", "ERROR", "WARNING", "NOTICE",
                                  "INFO", "DEBUG", "TRACE", "SSI"};
    std::fprintf(stderr, "[%s] %s\n", names[(level >= 1 && level <= 9) ? level : 0], message);
}

extern "C" void SoapySDR_log(const SoapySDRLogLevel logLevel, const char *message)
{
    if (logLevel > g_logLevel && logLevel != SOAPY_SDR_SSI)
        return;
    (g_logHandler ? g_logHandler : defaultLogHandler)(logLevel, message);
}

extern "C" void SoapySDR_vlogf(const SoapySDRLogLevel logLevel, const char *format,
                               va_list argList)
{
    if (logLevel > g_logLevel && logLevel != SOAPY_SDR_SSI)
        return;
    char buf[512];
    std::vsnprintf(buf, sizeof buf, format, argList);
    SoapySDR_log(logLevel, buf);
}

extern "C" void SoapySDR_logf(const SoapySDRLogLevel logLevel, const char *format, ...)
{
    if (logLevel > g_logLevel && logLevel != SOAPY_SDR_SSI)
        return;
    va_list args;
    va_start(args, format);
    SoapySDR_vlogf(logLevel, format, args);
    va_end(args);
}

extern "C" void SoapySDR_registerLogHandler(const SoapySDRLogHandler handler)
{
    g_logHandler = handler;
}
extern "C" void SoapySDR_setLogLevel(const SoapySDRLogLevel logLevel) { g_logLevel = logLevel; }
extern "C" SoapySDRLogLevel SoapySDR_getLogLevel(void) { return g_logLevel; }

void SoapySDR::log(const LogLevel logLevel, const std::string &message)
{
    SoapySDR_log(logLevel, message.c_str());
}
void SoapySDR::vlogf(const SoapySDRLogLevel logLevel, const char *format, va_list argList)
{
    SoapySDR_vlogf(logLevel, format, argList);
}
void SoapySDR::logf(const SoapySDRLogLevel logLevel, const char *format, ...)
{
    if (logLevel > g_logLevel && logLevel != SOAPY_SDR_SSI)
        return;
    va_list args;
    va_start(args, format);
    SoapySDR_vlogf(logLevel, format, args);
    va_end(args);
}
void SoapySDR::registerLogHandler(const LogHandler &handler) { g_logHandler = handler; }
void SoapySDR::setLogLevel(const LogLevel logLevel) { g_logLevel = logLevel; }
SoapySDR::LogLevel SoapySDR::getLogLevel(void) { return g_logLevel; }

// ---------------------------------------------------------------------------------------
// Tick/time conversion: one definition (csrc/sx_time.h) shared with the device code.
// ---------------------------------------------------------------------------------------
extern "C" long long SoapySDR_ticksToTimeNs(const long long ticks, const double rate)
{
    return sx_ticks_to_time_ns(ticks, rate);
}

extern "C" long long SoapySDR_timeNsToTicks(const long long timeNs, const double rate)
{
    return sx_time_ns_to_ticks(timeNs, rate);
}

extern "C" const char *SoapySDR_errToStr(int errorCode)
{
    switch (errorCode) {
    case SOAPY_SDR_TIMEOUT: return "TIMEOUT";
    case SOAPY_SDR_STREAM_ERROR: return "STREAM_ERROR";
    case SOAPY_SDR_CORRUPTION: return "CORRUPTION";
    case SOAPY_SDR_OVERFLOW: return "OVERFLOW";
    case SOAPY_SDR_NOT_SUPPORTED: return "NOT_SUPPORTED";
    case SOAPY_SDR_TIME_ERROR: return "TIME_ERROR";
    case SOAPY_SDR_UNDERFLOW: return "UNDERFLOW";
    default: return "UNKNOWN";
    }
}

extern "C" size_t SoapySDR_formatToSize(const char *format)
{
    // "C" prefix = complex (two components); digits = bits per component.
    size_t bits = 0;
    bool isComplex = false;
    for (const char *p = format; *p; p++) {
        if (*p == 'C')
            isComplex = true;
        if (*p >= '0' && *p <= '9')
            bits = bits * 10 + size_t(*p - '0');
    }
    return ((isComplex ? 2 : 1) * bits + 7) / 8;
}
