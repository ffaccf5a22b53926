/* stand-in for VTK's vtkSmartPointer.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
