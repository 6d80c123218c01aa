#pragma once
#include <mutex>
namespace PLPSLAM { namespace data {
class map_database {  // data/map_database.h
public:
    static std::mutex mtx_database_;
};
} }
