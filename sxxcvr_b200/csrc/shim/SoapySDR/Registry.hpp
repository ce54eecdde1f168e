// Minimal SoapySDR-compatible device registry (shim).  A driver module registers itself
// with a static Registry object, as the reference does at SoapySX.cpp:1656.
#pragma once
#include <SoapySDR/Types.hpp>
#include <SoapySDR/Version.h>
#include <map>
#include <string>
namespace SoapySDR {
class Device;
typedef KwargsList (*FindFunction)(const Kwargs &);
typedef Device *(*MakeFunction)(const Kwargs &);
typedef std::map<std::string, FindFunction> FindFunctions;
typedef std::map<std::string, MakeFunction> MakeFunctions;

class Registry {
public:
    Registry(const std::string &name, const FindFunction &find, const MakeFunction &make,
             const std::string &abi);
    ~Registry(void);
    static std::vector<std::string> listDrivers(void);
    static FindFunctions listFindFunctions(void);
    static MakeFunctions listMakeFunctions(void);

private:
    std::string _name;
};
}
