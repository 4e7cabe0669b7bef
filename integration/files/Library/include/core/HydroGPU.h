/*
    This file is an ADDITION to Stonefish (integration/stonefish_v1.6.0.patch of the stonefish_b200 repository).

    HydroGPU — batched evaluation of the geometry-based hydrodynamics on a CUDA device.

    Replaces, for all solids overlapping the ocean at once, what Ocean::ApplyFluidForces asks of every body when the forces are
    recomputed (Ocean.cpp:246-254): SolidEntity::ComputeHydrodynamicForces / Compound::ComputeHydrodynamicForces, i.e.
    CheckBodyFluidPosition, the Surface / Submerged integrators, the analytic buoyancy of fully submerged bodies and
    CorrectHydrodynamicForces. The results land in the same members (Fb, Tb, Fdq, Tdq, Fdf, Tdf, Swet, Vsub), members the
    CPU code would leave untouched stay untouched, and SolidEntity::ApplyHydrodynamicForces is used unchanged.
*/
#ifndef __Stonefish_HydroGPU__
#define __Stonefish_HydroGPU__

#include <vector>

class btCollisionObject;

namespace sf
{
    class Ocean;
    class SolidEntity;

    //! A class evaluating the hydrodynamic forces of a batch of solids on a GPU (libsfh.so).
    class HydroGPU
    {
    public:
        //! A constructor.
        /*!
         \param cudaDevice the index of the CUDA device to use
         */
        HydroGPU(int cudaDevice = 0);

        //! A destructor.
        ~HydroGPU();

        //! A method that recomputes the hydrodynamic forces of all passed solids (settings: damping forces and realistic buoyancy enabled).
        /*!
         \param ocn a pointer to the ocean
         \param solids the solids (also compounds) overlapping the ocean
         */
        void ComputeHydrodynamicForces(Ocean* ocn, const std::vector<SolidEntity*>& solids);

        //! A method that forgets all uploaded physics meshes (to be called when the scenario is destroyed).
        void Clear();

        //! A method returning the solid behind a collision object overlapping the ocean, or nullptr (static, kinematic, not a solid).
        static SolidEntity* SolidFromCollisionObject(btCollisionObject* co);

    private:
        struct Impl;
        Impl* impl;
    };
}

#endif
