from .base import (
    InfectionNetwork,
    InfectionNetworks,
    SchoolNetwork,
    CompanyNetwork,
    HouseholdNetwork,
    CareHomeNetwork,
    UniversityNetwork,
)
from .leisure_network import (
    LeisureNetwork,
    PubNetwork,
    GroceryNetwork,
    CinemaNetwork,
    VisitNetwork,
    GymNetwork,
    CareVisitNetwork,
)
