// WorldPhysicsLoader.hpp -- plugin entry with the shape of reference src/WorldPhysicsLoader.hpp:19-39.
#pragma once
#include "WorldPhysics.hpp"

namespace mars {
namespace ode_physics {

class WorldPhysicsLoader : public lib_manager::LibInterface {
public:
    explicit WorldPhysicsLoader(lib_manager::LibManager *theManager);
    virtual ~WorldPhysicsLoader(void);
    int getLibVersion() const override { return 1; }
    const std::string getLibName() const override { return std::string("mars_ode_physics"); }
    std::shared_ptr<interfaces::PhysicsInterface> createWorldInstance();
    // batch extension: every world created afterwards holds this many device replicas
    void setNumWorlds(int n) { num_worlds = n; }

private:
    int num_worlds = 1;
};

} // namespace ode_physics
} // namespace mars
