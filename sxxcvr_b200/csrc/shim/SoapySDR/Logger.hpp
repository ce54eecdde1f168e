// Minimal SoapySDR-compatible logger, C++ side (shim).  SoapySDR::logf is used at
// reference SoapySX.cpp:1649.
#pragma once
#include <SoapySDR/Logger.h>
#include <string>
namespace SoapySDR {
typedef SoapySDRLogLevel LogLevel;
typedef SoapySDRLogHandler LogHandler;
void log(const LogLevel logLevel, const std::string &message);
void vlogf(const SoapySDRLogLevel logLevel, const char *format, va_list argList);
void logf(const SoapySDRLogLevel logLevel, const char *format, ...);
void registerLogHandler(const LogHandler &handler);
void setLogLevel(const LogLevel logLevel);
LogLevel getLogLevel(void);
}
